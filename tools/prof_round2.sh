#!/bin/bash
# On the GPU box: the ncu evidence of round 2 (C2 workload).  Everything lands in gpurun_out/.
#   launch list of one bench step (cold-cache, serialised: compare SHARES), and --set full of the dominant kernels
mkdir -p gpurun_out
export FAUCET_BENCH_PROFILE_RUN=1 FAUCET_BENCH_SKIP_EXTRAS=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
for spec in "stitch_flow_kernel 1 flow" "load_A_kernel 40 loadA" "scan_flags_memo_kernel 1 scan" "radix_scatter_kernel 3 radix"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/r2_prof_$3 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_$3.log 2>&1
  ncu -i gpurun_out/r2_prof_$3.ncu-rep --page details > gpurun_out/r2_ncu_$3.txt 2>&1
done
ls -la gpurun_out/r2_*
