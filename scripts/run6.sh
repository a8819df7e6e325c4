mkdir -p gpurun_out
timeout 300 python scripts/flowtime.py > gpurun_out/flowtime_alu.log 2>&1
timeout 300 python scripts/flowprof.py > gpurun_out/flowprof_alu.log 2>&1
cat gpurun_out/flowtime_alu.log gpurun_out/flowprof_alu.log | head -5
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -s -k "c2 or C2" 2>&1 | grep -i "C2 full\|passed\|failed" 
timeout 1500 python -m pytest tests/test_gpu_discrete.py tests/test_gpu_fuzz.py -x -q > gpurun_out/t_discrete.log 2>&1; tail -4 gpurun_out/t_discrete.log
