set -x
python -m pytest tests/test_gpu_ops.py -q -x 2>&1 | tail -4
python tools/bench_bn.py 2>&1 | grep "'bn': 0" | cut -c1-200
VGEN_TAPGEMM_NARROW_STORE=1 python tools/bench_bn.py 2>&1 | grep "'bn': 0" | cut -c1-200
python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-decode > gpurun_out/r02m_bench_i2vgen.json 2>&1
VGEN_TAPGEMM_NARROW_STORE=1 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-decode > gpurun_out/r02m_bench_i2vgen_narrow.json 2>&1
