mkdir -p gpurun_out
timeout 600 python scripts/flow512_variants.py > gpurun_out/flow512_variants.log 2>&1; cat gpurun_out/flow512_variants.log
FB2_FLOW_DIRECT=0 timeout 1500 python -m pytest tests/test_gpu_discrete.py -x -q -k "flow or marginals or chunk_summary or time_sharded or host_entry" > gpurun_out/t_flow_twohop.log 2>&1; tail -5 gpurun_out/t_flow_twohop.log
