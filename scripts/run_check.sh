mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/t_gpu_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e']['ms_per_step'], d['extra']['kalman_scan']['ms'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -c 400
