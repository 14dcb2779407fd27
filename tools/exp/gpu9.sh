mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "retained or scan_matches" > gpurun_out/pytest_ret.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_ret.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_v10.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/bench_r1_v10.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value']/1e9, d['ms_per_step'], d['e2e'], d['kernels_ms_per_step'])
PY
tail -3 gpurun_out/bench_err.log
