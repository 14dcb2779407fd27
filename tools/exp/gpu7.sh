mkdir -p gpurun_out
timeout 900 python tools/stitch_sweep.py > gpurun_out/sweep7.log 2>&1; echo "sweep rc=$?"; tail -12 gpurun_out/sweep7.log
