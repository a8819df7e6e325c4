mkdir -p gpurun_out
FB2_FLOW_TC=1 timeout 300 python scripts/flowprof.py > gpurun_out/flowprof_tc.log 2>&1
cat gpurun_out/flowprof_tc.log
FB2_FLOW_TC=1 timeout 300 python scripts/flowtime.py > gpurun_out/flowtime_tc.log 2>&1
timeout 300 python scripts/flowtime.py > gpurun_out/flowtime.log 2>&1
tail -5 gpurun_out/flowtime_tc.log gpurun_out/flowtime.log
