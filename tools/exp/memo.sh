mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/pytest_memo.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_memo.log
for T in "memo_shift=1"; do
  FAUCET_TUNING=$T timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tmp_m.json 2> gpurun_out/bench_err.log
  python - "$T" <<'PY'
import json,sys
for l in open('gpurun_out/bench_tmp_m.json'):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], "%.3f G (%.1f ms) e2e %.3f G scan_flags %.2f ms" % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['kernels_ms_per_step']['scan_flags']))
PY
done
