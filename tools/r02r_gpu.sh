# r02r: op parity of everything new (attention variants incl. TS + FMA-pipe exp2, folded LayerNorm, row stats), attention timings,
# tiny-model parity + the config-2 full-size forward with the new defaults, and the config-2 step with / without the LN fold
set -x
timeout 700 python -m pytest tests/test_gpu_ops.py -q -x > gpurun_out/r02r_ops.log 2>&1; echo rc ops $?; tail -6 gpurun_out/r02r_ops.log | cut -c1-600
timeout 300 python tools/bench_attn.py --variants 2,t1,t1L,t1x4,t1x8,t2,t2x4,t2x8 > gpurun_out/r02r_attn.log 2>&1; echo rc attn $?
grep -E 'parity|"shape"|rror' gpurun_out/r02r_attn.log | cut -c1-175
timeout 500 python -m pytest tests/test_gpu_parity.py -q -x > gpurun_out/r02r_parity.log 2>&1; echo rc parity $?; tail -6 gpurun_out/r02r_parity.log | cut -c1-600
timeout 500 python -m pytest tests/test_gpu_fullsize.py -q -x -s -k "forward and full_i2vgen" > gpurun_out/r02r_full.log 2>&1; echo rc full $?; grep -E "fullsize\]|passed|failed|rror" gpurun_out/r02r_full.log | cut -c1-700
VGEN_LN_FOLD=1 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-decode > gpurun_out/r02r_bench_fold1.json 2> gpurun_out/r02r_bench_fold1.err; echo rc b1 $?; tail -c 300 gpurun_out/r02r_bench_fold1.err
VGEN_LN_FOLD=0 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-decode > gpurun_out/r02r_bench_fold0.json 2> gpurun_out/r02r_bench_fold0.err; echo rc b0 $?; tail -c 300 gpurun_out/r02r_bench_fold0.err
python - <<'PY'
import json
for f in ("gpurun_out/r02r_bench_fold1.json", "gpurun_out/r02r_bench_fold0.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), json.dumps(d.get("kernel_families"))[:600])
    except Exception as e:
        print(f, "unreadable", e)
PY
