timeout 240 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "2" > gpurun_out/r2s_pytest_mgpu2.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2s_pytest_mgpu2.log
bash tools/exp/t2.sh 2 "60"
