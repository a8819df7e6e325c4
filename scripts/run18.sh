mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gaussian.py tests/test_gpu_fullsize.py -x -q -s -k "kalman or c4 or C4 or gaussian or cta_wide" 2>&1 | grep -i "passed\|failed\|error\|assert" | tail -8
FB2_KALMAN_UPPER=0 timeout 300 python scripts/kalman_prof.py 20; timeout 300 python scripts/kalman_prof.py 20; FB2_KALMAN_K=4 timeout 300 python scripts/kalman_prof.py 20; FB2_KALMAN_K=6 timeout 300 python scripts/kalman_prof.py 20
