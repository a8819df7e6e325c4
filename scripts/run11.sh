mkdir -p gpurun_out
timeout 600 python scripts/flow512_variants.py > gpurun_out/flow512_variants2.log 2>&1; cat gpurun_out/flow512_variants2.log
timeout 600 python scripts/direct_time.py > gpurun_out/direct_time.log 2>&1; tail -8 gpurun_out/direct_time.log
timeout 1800 python -m pytest tests/test_gpu_discrete.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py -x -q -s 2>&1 | grep -v "^$" > gpurun_out/t_discrete.log; tail -8 gpurun_out/t_discrete.log
