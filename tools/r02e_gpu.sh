set -x
python -m pytest tests -m gpu -q -x > gpurun_out/r02e_tests.log 2>&1; tail -6 gpurun_out/r02e_tests.log
python tools/bench_attn.py > gpurun_out/r02e_attn.log 2>&1; tail -4 gpurun_out/r02e_attn.log | cut -c1-200
for w in i2vgen videolcm higen; do
  python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02e_bench_$w.json 2> gpurun_out/r02e_bench_$w.err; tail -c 300 gpurun_out/r02e_bench_$w.err
  VGEN_PDL=0 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-decode --full-video 0 --profile-pass 0 > gpurun_out/r02e_bench_${w}_nopdl.json 2>&1
done
