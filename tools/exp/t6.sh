export FAUCET_BENCH_SKIP_EXTRAS=1 FAUCET_BENCH_PROFILE_RUN=1
FAUCET_TUNING="epoch_mode=1,epoch0=900000,epoch_max=4000000,epoch_switch_pct=100" timeout 600 ncu --set full --clock-control none --import-source on -k regex:stitch_dry_kernel -s 0 -c 1 -o gpurun_out/r2s_prof_dry -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2s_ncu_dry.log 2>&1
echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stitch_flow_kernel -s 1 -c 1 -o gpurun_out/r2s_prof_flow -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2s_ncu_flow.log 2>&1
echo rc=$?
