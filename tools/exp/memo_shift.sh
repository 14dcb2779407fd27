# scan_flags with the memo at half its size (64 MiB at configs[1]: inside L2) against the default (128 MiB)
export FAUCET_BENCH_SKIP_EXTRAS=1
mkdir -p gpurun_out
for v in 2; do
  FAUCET_TUNING="memo_shift=$v" timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_memo_shift_$v.json 2> gpurun_out/r2_memo_shift_${v}_err.log; echo "rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_memo_shift_$v.json").read().strip().splitlines()[-1])
print("memo_shift=$v", round(d["ms_per_step"],1), "ms", {k:round(x,2) for k,x in d["kernels_ms_per_step"].items() if x})
PY
done
