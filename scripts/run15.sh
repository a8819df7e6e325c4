mkdir -p gpurun_out
timeout 600 python scripts/kalman_tc_time.py > gpurun_out/kalman_tc_time.log 2>&1; cat gpurun_out/kalman_tc_time.log
timeout 900 python -m pytest tests/test_gpu_gaussian.py -x -q -k "kalman" > gpurun_out/t_kalman.log 2>&1; tail -12 gpurun_out/t_kalman.log
