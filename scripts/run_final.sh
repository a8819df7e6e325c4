# evidence run: full GPU suite, bench line, launch lists, ncu captures (each profiler run only after its command passed without ncu)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/t_gpu_all.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_gpu_all.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_n1.json
timeout 600 python scripts/kalman_prof.py 5 > gpurun_out/kalman_prof.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gaussian -c 40 --csv --log-file gpurun_out/r02_kalman_final_launches.csv python scripts/kalman_prof.py 1 > gpurun_out/ncu_kalman.log 2>&1
timeout 300 python scripts/ncu_probe.py > gpurun_out/ncu_probe_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hmm_flow_forward -c 1 -o gpurun_out/r02_flow_final_full python scripts/ncu_probe.py > gpurun_out/ncu_flow.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_bench_final.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 300 python scripts/flowprof.py > gpurun_out/flowprof_final.log 2>&1; head -3 gpurun_out/flowprof_final.log
ls -la gpurun_out | tail -8
