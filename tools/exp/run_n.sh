mkdir -p gpurun_out
N=$1
if [ "$2" = "test" ]; then timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_mgpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_mgpu.log; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_r1_v15_n$N.json 2> gpurun_out/bench_n${N}_err.log; echo "bench$N rc=$?"
python - $N <<'PY'
import json,sys
for l in open('gpurun_out/bench_r1_v15_n%s.json' % sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print("N=%s: %.3f G k-mers/s (%.1f ms), e2e %.3f G" % (sys.argv[1], d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9)); print(d['kernels_ms_per_step'])
PY
tail -2 gpurun_out/bench_n${N}_err.log | cut -c1-200
