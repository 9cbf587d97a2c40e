# r02t: the last GPU seconds of the round: parity (9 small cases vs fp32 SDPA) and timing of the three-CTAs-per-SM TS instantiation
timeout 60 python tools/bench_attn.py --only 0,2,4 --variants t3,2,t1 > gpurun_out/r02t_attn.log 2>&1; echo rc $?
grep -E 'parity|"shape"|rror|timed out' gpurun_out/r02t_attn.log | cut -c1-330
