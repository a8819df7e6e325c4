mkdir -p gpurun_out
timeout 300 python scripts/flowtime.py > gpurun_out/flowtime_f.log 2>&1
timeout 300 python scripts/flowprof.py > gpurun_out/flowprof_f.log 2>&1
cat gpurun_out/flowtime_f.log gpurun_out/flowprof_f.log | head -4
timeout 1500 python -m pytest tests/test_gpu_discrete.py -x -q -k "flow or marginals or both_dataflow or chunk_summary or time_sharded or host_entry" > gpurun_out/t_flow.log 2>&1; tail -12 gpurun_out/t_flow.log
