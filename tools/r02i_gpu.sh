set -x
python -m pytest tests -m gpu -q > gpurun_out/r02i_tests.log 2>&1; tail -4 gpurun_out/r02i_tests.log | cut -c1-300
for w in i2vgen videolcm higen sr600; do
  python bench.py --workload $w --steps 8 --warmup 3 > gpurun_out/r02i_bench_$w.json 2> gpurun_out/r02i_bench_$w.err; tail -c 300 gpurun_out/r02i_bench_$w.err
done
cp gpurun_out/prof_shapes_i2vgen.json gpurun_out/r02i_shapes_i2vgen.json
VGEN_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2800 --csv --log-file gpurun_out/launches_r02i.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-decode --no-e2e --profile-pass 0 > gpurun_out/r02i_ncu_bench.log 2>&1
ls -la gpurun_out | grep r02i
