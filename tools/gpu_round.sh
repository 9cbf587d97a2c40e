#!/bin/bash
# One GPU session: gate tests, the bench line, a launch list and one ncu --set full capture of the
# two top kernels.  Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?"; tail -c 5000 gpurun_out/bench_$TAG.log
if [ "$2" != "noprof" ]; then
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1320 -c 1316 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-decode --profile-pass 0 > gpurun_out/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tapgemm_sm100 -s 300 -c 3 -o gpurun_out/prof_tapgemm_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-decode --profile-pass 0 > gpurun_out/ncu_tap_$TAG.log 2>&1; echo "ncu tapgemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_sm100 -s 20 -c 2 -o gpurun_out/prof_attn_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-decode --profile-pass 0 > gpurun_out/ncu_attn_$TAG.log 2>&1; echo "ncu attn rc=$?"
fi
ls -la gpurun_out | tail -12
