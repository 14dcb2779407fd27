export FAUCET_BENCH_SKIP_EXTRAS=1 FAUCET_BENCH_PROFILE_RUN=1
for v in "dry_lazy=0" "dry_lazy=1"; do
FAUCET_TUNING="epoch_mode=1,epoch0=900000,epoch_max=4000000,epoch_switch_pct=100,$v" timeout 600 ncu --set full --clock-control none --import-source on -k regex:stitch_dry_kernel -s 0 -c 1 -o gpurun_out/r2s_prof_dry_${v#dry_lazy=} -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2s_ncu_dry_${v#dry_lazy=}.log 2>&1
echo rc=$?
done
ls -la gpurun_out/r2s_prof_dry*
