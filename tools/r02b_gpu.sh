set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -q -x > gpurun_out/r02b_tests.log 2>&1; tail -3 gpurun_out/r02b_tests.log
python bench.py --steps 8 --warmup 3 > gpurun_out/r02b_bench_i2vgen.json 2> gpurun_out/r02b_bench_i2vgen.err; tail -c 600 gpurun_out/r02b_bench_i2vgen.err
cp gpurun_out/prof_shapes_i2vgen.json gpurun_out/r02b_shapes_i2vgen.json
VGEN_CUDA_GRAPH=0 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-decode --profile-pass 0 > gpurun_out/r02b_bench_i2vgen_nograph.json 2>&1
for w in videolcm higen sr600; do
  python bench.py --workload $w --steps 8 --warmup 3 > gpurun_out/r02b_bench_$w.json 2> gpurun_out/r02b_bench_$w.err; tail -c 400 gpurun_out/r02b_bench_$w.err
done
VGEN_CUDA_GRAPH=0 python bench.py --workload videolcm --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-decode --profile-pass 0 > gpurun_out/r02b_bench_videolcm_nograph.json 2>&1
VGEN_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2800 --csv --log-file gpurun_out/launches_r02b.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-decode --no-e2e --profile-pass 0 > gpurun_out/r02b_ncu_bench.log 2>&1
ls -la gpurun_out | tail -15
