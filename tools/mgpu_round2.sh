#!/bin/bash
# On an N-GPU box (gpurun --gpus 8 -- 'bash tools/mgpu_round2.sh'): sharded parity tests at 2/4/8 GPUs, then the bench at N = 8, 4, 2.
mkdir -p gpurun_out
export FAUCET_BENCH_SKIP_EXTRAS=1
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2_pytest_mgpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_mgpu.log
NG=$(nvidia-smi -L | wc -l)
for n in 8 4 2; do
  [ $n -le $NG ] || continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n${n}_err.log; echo "bench n=$n rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_n$n.json").read().strip().splitlines()[-1])
    print("N=$n", round(d["value"]/1e9,3), "G k-mers/s", round(d["ms_per_step"],1), "ms; e2e", round(d["e2e"]["value"]/1e9,3), d.get("parity_check"), {k:round(v,1) for k,v in d["kernels_ms_per_step"].items() if v})
except Exception as e: print("N=$n failed", e)
PY
done
if [ 8 -le $NG ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --workload c3 --scaling strong --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3_strong_n8.json 2> gpurun_out/r2_bench_c3_strong_n8_err.log; echo "c3 strong rc=$?"; tail -c 1500 gpurun_out/r2_bench_c3_strong_n8.json
fi
