mkdir -p gpurun_out
timeout 900 python tools/stitch_sweep.py > gpurun_out/sweep_rw2.log 2>&1; echo "sweep rc=$?"; tail -4 gpurun_out/sweep_rw2.log | cut -c1-600
