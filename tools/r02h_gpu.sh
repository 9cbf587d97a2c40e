set -x
python tools/prof_decode.py > gpurun_out/r02h_decode.log 2>&1; cat gpurun_out/r02h_decode.log | cut -c1-200
for p in 0 8 4 2; do VGEN_ATTN_POLY=$p python tools/bench_attn.py --only 0 2>&1 | grep '"shape"' | cut -c1-120; done
for p in 0 4; do VGEN_ATTN_POLY=$p python tools/bench_attn.py --only 1 2>&1 | grep '"shape"' | cut -c1-120; done
python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -q -x -k "clip or attention or variants or vae" 2>&1 | tail -5
