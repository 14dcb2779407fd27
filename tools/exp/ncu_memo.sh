mkdir -p gpurun_out
SWEEP='{"stitch_impl":1}' timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_flags_memo -c 1 -o gpurun_out/prof_scan_memo_v14 python tools/stitch_sweep.py > gpurun_out/ncu_scan_memo.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_scan_memo.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v14.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "launchlist rc=$?"
