for m in 64 256 512 1024 2048 100000; do echo "upper_max=$m"; FB2_KALMAN_UPPER_MAX=$m timeout 300 python scripts/kalman_prof.py 20; done
FB2_GUARD=1 timeout 900 python -m pytest tests/test_gpu_gaussian.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -3
