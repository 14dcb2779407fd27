# single GPU: one ordered epoch of 900 k records, then classify epochs: time of the read-only walk per record, variants
export FAUCET_BENCH_SKIP_EXTRAS=1
for v in "dry_lazy=0" "dry_lazy=1" $EXTRA; do
  FAUCET_TUNING="epoch_mode=1,epoch0=900000,epoch_max=4000000,epoch_switch_pct=100,$v" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_dry_$v.json 2> gpurun_out/r2s_dry_${v}_err.log; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s_dry_$v.json").read().strip().splitlines()[-1])
    print("$v", round(d["ms_per_step"],1), "ms", {k:round(x,2) for k,x in d["kernels_ms_per_step"].items() if x}, d.get("stitch"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2s_dry_${v}_err.log").read()[-2000:])
PY
done
