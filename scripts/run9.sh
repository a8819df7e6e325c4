mkdir -p gpurun_out
timeout 600 python scripts/kalman_fused_time.py > gpurun_out/kalman_fused_time2.log 2>&1; cat gpurun_out/kalman_fused_time2.log
timeout 900 python -m pytest tests/test_gpu_gaussian.py -x -q -k kalman > gpurun_out/t_kalman.log 2>&1; tail -3 gpurun_out/t_kalman.log
