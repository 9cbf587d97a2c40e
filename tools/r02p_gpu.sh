# r02p: three-q-tile / 64-key-block attention instantiation: op parity (both variants), then timings of every config-2 shape
set -x
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "attention" > gpurun_out/r02p_attn_tests.log 2>&1; echo rc $?; tail -15 gpurun_out/r02p_attn_tests.log | cut -c1-400
timeout 600 python tools/bench_attn.py > gpurun_out/r02p_attn.log 2>&1; echo rc $?; grep '"shape"' gpurun_out/r02p_attn.log | cut -c1-420
