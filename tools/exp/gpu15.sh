mkdir -p gpurun_out
SWEEP='{"stitch_impl":1}' timeout 900 ncu --set full --clock-control none --import-source on -k regex:load_A -s 12 -c 1 -o gpurun_out/prof_loadA_v12 python tools/stitch_sweep.py > gpurun_out/ncu_loadA.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_loadA.log
