mkdir -p gpurun_out
for WT in "c3:stitch_impl=1" "c3:stitch_impl=2" "c4:stitch_impl=2" "c1:stitch_impl=2"; do
  W=${WT%%:*}; T=${WT##*:}
  FAUCET_TUNING=$T timeout 2400 python bench.py --steps 2 --warmup 3 --workload $W --no-cpu-baseline > gpurun_out/bench_tmp_x.json 2> gpurun_out/bench_err_x.log; echo "$WT rc=$?"
  python - <<'PY'
import json,sys
for l in open('gpurun_out/bench_tmp_x.json'):
    if l.startswith('{'):
        d=json.loads(l); print("%.3f G k-mers/s resident (%.1f ms), e2e %.3f G, junctions %d" % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['config']['junctions'])); print(d['kernels_ms_per_step'], d['stitch']['stitch_rounds'])
PY
  tail -2 gpurun_out/bench_err_x.log
done
