mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_mgpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_mgpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_r1_v9_n8.json 2> gpurun_out/bench_n8_err.log; echo "bench8 rc=$?"; tail -c 1800 gpurun_out/bench_r1_v9_n8.json; tail -5 gpurun_out/bench_n8_err.log
