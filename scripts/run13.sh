mkdir -p gpurun_out
FB2_FLOW_DIRECT=0 timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -s -k "c5" 2>&1 | grep -i "full size\|passed\|failed\|assert" | head
timeout 600 python scripts/full_configs.py 2>&1 | tail -8
