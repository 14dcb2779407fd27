#!/bin/bash
# On the GPU box: ncu --set full of the stitch kernels on a bench workload (one launch each, written as .ncu-rep + text details)
#   bash tools/prof_stitch.sh <tag> [workload]
tag=${1:-r2}; wl=${2:-c2}
mkdir -p gpurun_out
# the dry (classify) kernel: take a late, large launch; the dataflow executor: epoch_mode 0 gives one big launch
SWEEP='{}' timeout 900 ncu --set full --clock-control none --import-source on -k regex:stitch_dry_kernel -s 6 -c 1 \
  -o gpurun_out/prof_dry_$tag -f python tools/stitch_sweep.py $wl > gpurun_out/ncu_dry_$tag.log 2>&1
SWEEP='{"epoch_mode":0}' timeout 900 ncu --set full --clock-control none --import-source on -k regex:stitch_flow_kernel -s 0 -c 1 \
  -o gpurun_out/prof_flow_$tag -f python tools/stitch_sweep.py $wl > gpurun_out/ncu_flow_$tag.log 2>&1
ls -la gpurun_out/prof_*_$tag.ncu-rep
