# r02v: last call of the round: the final library (rebuilt after r02s's full gate run) through the model-level parity tests + smoke()
timeout 70 python -m pytest tests/test_gpu_parity.py -q -x > gpurun_out/r02v_parity.log 2>&1; echo rc $?; tail -3 gpurun_out/r02v_parity.log | cut -c1-300
