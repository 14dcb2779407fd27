#!/bin/bash
# On the GPU box: ncu --set full of one stitch kernel launch on a bench workload (.ncu-rep into gpurun_out/)
#   bash tools/prof_stitch.sh <tag> <kernel regex> <launch skip> '<SWEEP json>' [workload]
tag=$1; kern=$2; skip=$3; cfg=$4; wl=${5:-c2}
mkdir -p gpurun_out
SWEEP="$cfg" timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 \
  -o gpurun_out/prof_$tag -f python tools/stitch_sweep.py $wl > gpurun_out/ncu_$tag.log 2>&1
ls -la gpurun_out/prof_$tag.ncu-rep
