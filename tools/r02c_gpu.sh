set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_video.py -q > gpurun_out/r02c_tests.log 2>&1; tail -5 gpurun_out/r02c_tests.log
python bench.py --workload higen --steps 8 --warmup 3 > gpurun_out/r02c_bench_higen.json 2> gpurun_out/r02c_bench_higen.err; tail -c 600 gpurun_out/r02c_bench_higen.err
python tools/bench_attn.py > gpurun_out/r02c_attn.log 2>&1; tail -20 gpurun_out/r02c_attn.log
VGEN_GN_SILU=1 python tools/prof_gn.py > gpurun_out/r02c_gn_mode1.log 2>&1
VGEN_GN_SILU=2 python tools/prof_gn.py > gpurun_out/r02c_gn_mode2.log 2>&1
paste -d'\n' gpurun_out/r02c_gn_mode1.log gpurun_out/r02c_gn_mode2.log | head -14
VGEN_GN_SILU=2 python -m pytest tests/test_gpu_ops.py -q -k "norm" 2>&1 | tail -3
