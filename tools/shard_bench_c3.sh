# On an N-GPU box (N >= 4: a shard must stay below 3 GiB of text): gpurun --gpus N -- bash tools/shard_bench_c3.sh N
# configs[2] (64 Mbp, 50x, 100 bp, k = 27) split over the GPUs (strong scaling).
mkdir -p gpurun_out
export FAUCET_BENCH_SKIP_EXTRAS=1
N=$1; shift
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus $N --workload c3 --scaling strong --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r2s_bench_c3_strong_n$N.json 2> gpurun_out/r2s_bench_c3_strong_n${N}_err.log; echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s_bench_c3_strong_n$N.json").read().strip().splitlines()[-1])
    print("c3 strong N=$N", round(d["value"]/1e9,3), "G k-mers/s", round(d["ms_per_step"],1), "ms; e2e", round(d["e2e"]["value"]/1e9,3), d.get("parity_check"), {k:round(v,1) for k,v in d["kernels_ms_per_step"].items() if v}, d.get("stitch_across_gpus"), d.get("stitch"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2s_bench_c3_strong_n${N}_err.log").read()[-3000:])
PY
