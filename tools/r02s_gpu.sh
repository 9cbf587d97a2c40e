# r02s: final-code run of round 2: the -m gpu gate, the bench lines, the ncu launch list of the bench command, the attention
# variants table, compute-sanitizer over the new kernels (TS attention, folded-LayerNorm epilogues), --set full captures.
# Ordered by priority; everything under its own timeout.
set -x
timeout 1000 python -m pytest tests -m gpu -q > gpurun_out/r02s_tests.log 2>&1; echo rc tests $?; tail -5 gpurun_out/r02s_tests.log | cut -c1-400
timeout 500 python bench.py --steps 8 --warmup 3 > gpurun_out/r02s_bench_i2vgen.json 2> gpurun_out/r02s_bench_i2vgen.err; echo rc bench $?; tail -c 300 gpurun_out/r02s_bench_i2vgen.err
cp gpurun_out/prof_shapes_i2vgen.json gpurun_out/r02s_shapes_i2vgen.json
VGEN_CUDA_GRAPH=0 timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2800 --csv --log-file gpurun_out/launches_r02s.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-decode --no-e2e --profile-pass 0 > gpurun_out/r02s_ncu_bench.log 2>&1; echo rc ncu-list $?
for w in videolcm higen; do
  timeout 400 python bench.py --workload $w --steps 8 --warmup 3 > gpurun_out/r02s_bench_$w.json 2> gpurun_out/r02s_bench_$w.err; echo rc bench $w $?; tail -c 200 gpurun_out/r02s_bench_$w.err
done
timeout 200 python tools/bench_attn.py --variants auto,2,3,t1,t2 > gpurun_out/r02s_attn.log 2>&1; echo rc attn $?; grep -E 'parity|"shape"' gpurun_out/r02s_attn.log | cut -c1-150
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
VGEN_CHECK_SMALL=1 timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_ops.py -q -x -k "attention or tapgemm_2cta or tapgemm_1cta or norm" > gpurun_out/r02s_memcheck.log 2>&1; echo memcheck rc $?; tail -4 gpurun_out/r02s_memcheck.log | cut -c1-300
VGEN_CHECK_SMALL=1 timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_ops.py -q -x -k "attention" > gpurun_out/r02s_racecheck.log 2>&1; echo racecheck rc $?; tail -4 gpurun_out/r02s_racecheck.log | cut -c1-300
N="ncu --set full --clock-control none --import-source on"
timeout 200 $N -k regex:attn_sm100 -s 10 -c 1 -f -o gpurun_out/prof_r02s_attn_ts1_cross python tools/bench_attn.py --only 4 --variants t1 > /dev/null 2>&1; echo rc $?
timeout 200 $N -k regex:attn_sm100 -s 10 -c 1 -f -o gpurun_out/prof_r02s_attn_ts1_880 python tools/bench_attn.py --only 2 --variants t1 > /dev/null 2>&1; echo rc $?
timeout 200 $N -k regex:attn_sm100 -s 10 -c 1 -f -o gpurun_out/prof_r02s_attn_ts1_14080 python tools/bench_attn.py --only 0 --variants t1 > /dev/null 2>&1; echo rc $?
timeout 200 $N -k regex:attn_sm100 -s 10 -c 1 -f -o gpurun_out/prof_r02s_attn_ss2_14080 python tools/bench_attn.py --only 0 --variants 2 > /dev/null 2>&1; echo rc $?
timeout 400 python bench.py --workload sr600 --steps 8 --warmup 3 > gpurun_out/r02s_bench_sr600.json 2> gpurun_out/r02s_bench_sr600.err; echo rc bench sr600 $?; tail -c 200 gpurun_out/r02s_bench_sr600.err
ls -la gpurun_out | grep r02s
