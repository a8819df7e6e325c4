mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gaussian.py tests/test_gpu_fullsize.py -x -q -s -k "kalman or c4 or C4 or gaussian" 2>&1 | grep -i "passed\|failed\|error\|C4" | tail -8
timeout 300 python scripts/kalman_prof.py 10 > gpurun_out/kalman_prof.log 2>&1; cat gpurun_out/kalman_prof.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gaussian -c 60 --csv --log-file gpurun_out/r02_kalman_tc_launches.csv python scripts/kalman_prof.py 1 > gpurun_out/ncu_kalman.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gaussian_kalman_tile_tc -c 1 -o gpurun_out/r02_kalman_tile_tc_full python scripts/kalman_prof.py 1 > gpurun_out/ncu_ktile.log 2>&1
ls -la gpurun_out/r02_kalman_t*
