mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_all.log
python - <<'PY'
import sys, time, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, bench, faucet_b200 as fb
w = bench.WORKLOADS["c2"]; k = w["k"]
_, lt, nh = fb.geometry_from_reads(w["est"], w["sing"], bench.FP)
path = bench.gen_dataset(w, seed=1)
for rep in range(3):
    t0 = time.perf_counter(); b2, _, st = fb.load_two_filters(path, True, k, lt, nh); t1 = time.perf_counter()
    recs, sst = fb.scan(path, True, True, True, k, bench.J, bench.MAX_SPACER, b2, lt, nh); t2 = time.perf_counter()
    print("file API: load %.1f ms scan %.1f ms  (%d k-mers, %d junctions)" % (1e3*(t1-t0), 1e3*(t2-t1), st.kmers, len(recs)))
PY
