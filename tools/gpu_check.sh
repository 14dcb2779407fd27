#!/bin/bash
# On the GPU box (gpurun -- 'bash tools/gpu_check.sh [bench args]'): the whole gpu-marked test-suite, then the bench.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/pytest_gpu_all.log
timeout 900 python bench.py --steps 5 --warmup 3 "$@" > gpurun_out/bench_last.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_last.json'):
    if l.startswith('{'):
        d = json.loads(l)
        r = d['roofline']
        print("resident %.3f G k-mers/s (%.1f ms)  e2e %.3f G  roofline: %s %.3f of peak (share of step %.2f), survey path %.3f" % (
            d['value'] / 1e9, d['ms_per_step'], d['e2e']['value'] / 1e9, r['kernel'], r['frac'], r['share_of_step'], r['survey_path_frac']))
        print(d['kernels_ms_per_step'], d['e2e'], d.get('e2e_file'), d.get('e2e_cleaning'), d.get('cpu_baseline'))
PY
tail -3 gpurun_out/bench_err.log
