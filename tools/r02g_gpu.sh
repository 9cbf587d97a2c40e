set -x
python -m pytest tests -m gpu -q > gpurun_out/r02g_tests.log 2>&1; tail -8 gpurun_out/r02g_tests.log | cut -c1-300
python tools/bench_attn.py > gpurun_out/r02g_attn.log 2>&1; tail -3 gpurun_out/r02g_attn.log | cut -c1-200
python bench.py --steps 8 --warmup 3 > gpurun_out/r02g_bench_i2vgen.json 2> gpurun_out/r02g_bench_i2vgen.err; tail -c 300 gpurun_out/r02g_bench_i2vgen.err
bash tools/ncu_capture.sh > gpurun_out/r02g_ncu_capture.log 2>&1; tail -12 gpurun_out/r02g_ncu_capture.log
