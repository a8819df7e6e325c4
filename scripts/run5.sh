mkdir -p gpurun_out
timeout 300 python scripts/flowtime.py > gpurun_out/flowtime_bf.log 2>&1
timeout 300 python scripts/flowprof.py > gpurun_out/flowprof_bf.log 2>&1
cat gpurun_out/flowtime_bf.log gpurun_out/flowprof_bf.log | head -6
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -k "c2 or C2" > gpurun_out/t_full.log 2>&1; tail -5 gpurun_out/t_full.log
timeout 1500 python -m pytest tests/test_gpu_discrete.py tests/test_gpu_fuzz.py -x -q > gpurun_out/t_discrete.log 2>&1; tail -15 gpurun_out/t_discrete.log
