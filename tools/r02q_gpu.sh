# r02q: TS-form attention family (P in tensor memory): each variant in its own process (a trap would poison the context):
# parity on small cases vs fp32 SDPA, then timings of the config-2 shapes; the SS-form kernels without the redundant P wait
set -x
for v in t1 t1p t2 t2p; do
  timeout 200 python tools/bench_attn.py --variants $v > gpurun_out/r02q_attn_$v.log 2>&1; echo rc $v $?
  grep -E 'parity|"shape"|rror|timed out' gpurun_out/r02q_attn_$v.log | cut -c1-330
done
timeout 200 python tools/bench_attn.py --variants 2,2n,3,3n > gpurun_out/r02q_attn_ss.log 2>&1; echo rc ss $?
grep -E 'parity|"shape"|rror' gpurun_out/r02q_attn_ss.log | cut -c1-330
