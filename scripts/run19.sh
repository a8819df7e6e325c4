mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_gaussian.py tests/test_gpu_fullsize.py tests/test_gpu_plugin.py tests/test_gpu_funsor.py tests/test_gpu_hmm_classes.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python scripts/kalman_prof.py 20
