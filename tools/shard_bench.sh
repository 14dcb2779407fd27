# On an N-GPU box: gpurun --gpus N -- bash tools/shard_bench.sh N "pct pct ..." [extra bench args]
# bench.py at N GPUs, once per listed share of shard 0 that GPU 0 stitches in order before the sharded epoch (FAUCET_SHARD=0
# in the environment: the serial stitch on GPU 0); one summary line per run, JSON lines in gpurun_out/.
mkdir -p gpurun_out
N=$1; PCTS=$2; shift 2
export FAUCET_BENCH_SKIP_EXTRAS=1
for pct in $PCTS; do
  FAUCET_SHARD_PREFIX_PCT=$pct timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+pct)) bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r2s_bench_n${N}_p$pct.json 2> gpurun_out/r2s_bench_n${N}_p${pct}_err.log; echo "bench n=$N pct=$pct rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s_bench_n${N}_p$pct.json").read().strip().splitlines()[-1])
    print("N=$N pct=$pct", round(d["value"]/1e9,3), "G k-mers/s", round(d["ms_per_step"],1), "ms; e2e", round(d["e2e"]["value"]/1e9,3), d.get("parity_check"), {k:round(v,1) for k,v in d["kernels_ms_per_step"].items() if v}, d.get("stitch_across_gpus"), d.get("stitch"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2s_bench_n${N}_p${pct}_err.log").read()[-3000:])
PY
done
