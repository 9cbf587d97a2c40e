set -x
for a in 0 1 0 1; do for s in 0 1; do VGEN_ATTN_ALT=$a python tools/bench_attn.py --only $s 2>&1 | grep '"shape"' | cut -c1-330; done; done
VGEN_ATTN_ALT=1 python -m pytest tests/test_gpu_ops.py -q -x -k "attention" 2>&1 | tail -2
VGEN_ATTN_ALT=1 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-decode > gpurun_out/r02k_bench_i2vgen_alt.json 2>&1
VGEN_ATTN_ALT=0 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-decode > gpurun_out/r02k_bench_i2vgen_noalt.json 2>&1
