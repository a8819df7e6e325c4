mkdir -p gpurun_out
FB2_CG_MODE=8 timeout 300 python scripts/summary_time.py 512 32768 2>&1 | tail -1
timeout 300 python scripts/summary_time.py 512 32768 2>&1 | tail -1
timeout 300 python scripts/summary_time.py 1024 8192 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_chain_gemm.py -x -q 2>&1 | tail -5
