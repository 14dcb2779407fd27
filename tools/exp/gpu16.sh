mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "load" > gpurun_out/pytest_load.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_load.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tmp.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/bench_tmp.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e'].get('load_call_ms'), d['e2e'].get('scan_call_ms'), d['kernels_ms_per_step'])
PY
tail -3 gpurun_out/bench_err.log
