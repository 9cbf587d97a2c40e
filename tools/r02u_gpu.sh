# r02u: op parity of every attention instantiation incl. TS3 (the -m gpu gate's attention group), TS3 on the remaining shapes
timeout 100 python -m pytest tests/test_gpu_ops.py -q -x -k attention > gpurun_out/r02u_attn_tests.log 2>&1; echo rc $?; tail -4 gpurun_out/r02u_attn_tests.log | cut -c1-500
timeout 40 python tools/bench_attn.py --only 1,3,5 --variants t3,t1 > gpurun_out/r02u_attn.log 2>&1; echo rc $?
grep -E 'parity|"shape"|rror' gpurun_out/r02u_attn.log | cut -c1-200
