set -x
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_ops.py tests/test_gpu_video.py -q -x -k "attention or variants or elementwise or video or frames" > gpurun_out/r02o_memcheck.log 2>&1; echo memcheck rc $?; tail -4 gpurun_out/r02o_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_ops.py -q -x -k "attention" > gpurun_out/r02o_racecheck.log 2>&1; echo racecheck rc $?; tail -4 gpurun_out/r02o_racecheck.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
