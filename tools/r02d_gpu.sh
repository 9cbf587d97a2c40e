set -x
python tools/bench_attn.py > gpurun_out/r02d_attn.log 2>&1; grep -c shape gpurun_out/r02d_attn.log; tail -3 gpurun_out/r02d_attn.log | cut -c1-400
for w in i2vgen videolcm; do
  python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-decode --full-video 0 > gpurun_out/r02d_bench_$w.json 2> gpurun_out/r02d_bench_$w.err; tail -c 300 gpurun_out/r02d_bench_$w.err
done
VGEN_ATTN_STAGGER=3 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-decode > gpurun_out/r02d_bench_i2vgen_stag3.json 2>&1
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 30 python -m pytest tests/test_gpu_ops.py -q -x > gpurun_out/r02d_memcheck.log 2>&1; echo memcheck rc $?; tail -5 gpurun_out/r02d_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 30 python -m pytest tests/test_gpu_ops.py -q -x -k "norm or attention or tapgemm_2cta" > gpurun_out/r02d_racecheck.log 2>&1; echo racecheck rc $?; tail -5 gpurun_out/r02d_racecheck.log
