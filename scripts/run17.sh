mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gaussian.py tests/test_gpu_fullsize.py -x -q -s -k "kalman or c4 or C4 or gaussian or cta_wide" 2>&1 | grep -i "passed\|failed\|error\|C4\|assert" | tail -8
timeout 300 python scripts/kalman_prof.py 10; FB2_GAUSS_LEVELS_WARP=1 timeout 300 python scripts/kalman_prof.py 10
FB2_KALMAN_K=6 timeout 300 python scripts/kalman_prof.py 10; FB2_KALMAN_K=4 timeout 300 python scripts/kalman_prof.py 10
