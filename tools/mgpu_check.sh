#!/bin/bash
# On an N-GPU box (gpurun --gpus N -- bash tools/mgpu_check.sh N): the sharded parity test at N GPUs, then the bench at N.
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "$N" > gpurun_out/pytest_mgpu$N.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_mgpu$N.log
bash tools/shard_bench.sh $N "${2:-70}"
