set -x
mkdir -p gpurun_out
timeout 120 ./scripts/microbench/probe > gpurun_out/probe.log 2>&1
timeout 300 python scripts/flowprof.py > gpurun_out/flowprof.log 2>&1
timeout 900 python -m pytest tests/test_gpu_gaussian.py -x -q -k "kalman" > gpurun_out/t_kalman.log 2>&1; tail -5 gpurun_out/t_kalman.log
timeout 300 python scripts/kalman_fused_time.py > gpurun_out/kalman_fused_time.log 2>&1
cat gpurun_out/probe.log gpurun_out/flowprof.log gpurun_out/kalman_fused_time.log
