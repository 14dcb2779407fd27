mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest rc=$?"; tail -16 gpurun_out/pytest_gpu_all.log
