timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "8" > gpurun_out/r2s_pytest_mgpu8.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2s_pytest_mgpu8.log
bash tools/exp/t2.sh 8 "70"
