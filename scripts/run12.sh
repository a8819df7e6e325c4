mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_discrete.py -x -q -s -k "full_size or flow or marginals or both_dataflow or chunk or time_sharded or host_entry" 2>&1 | grep -v "^$" > gpurun_out/t_sel.log; grep -i "full size\|passed\|failed" gpurun_out/t_sel.log | tail -12
timeout 600 python scripts/flow512_variants.py > gpurun_out/flow512_variants3.log 2>&1; head -3 gpurun_out/flow512_variants3.log
