mkdir -p gpurun_out
timeout 600 python scripts/viterbi_variants.py > gpurun_out/viterbi_variants.log 2>&1; cat gpurun_out/viterbi_variants.log
timeout 1200 python -m pytest tests/test_gpu_discrete.py -x -q -k "viterbi" > gpurun_out/t_viterbi.log 2>&1; tail -4 gpurun_out/t_viterbi.log
