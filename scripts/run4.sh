mkdir -p gpurun_out
for m in 1 33 0 32; do FB2_FLOW_MODE=$m timeout 300 python scripts/flowtime.py 2>&1 | tail -1; done > gpurun_out/flowtime_modes.log
FB2_FLOW_MODE=33 timeout 300 python scripts/flowprof.py > gpurun_out/flowprof_m33.log 2>&1
cat gpurun_out/flowtime_modes.log gpurun_out/flowprof_m33.log
