mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/pytest_rw.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_rw.log
timeout 900 python tools/stitch_sweep.py > gpurun_out/sweep_rw.log 2>&1; echo "sweep rc=$?"; tail -9 gpurun_out/sweep_rw.log
