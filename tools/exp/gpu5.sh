mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_v9.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_r1_v9.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v9.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1; echo "launchlist rc=$?"
SWEEP='{"stitch_impl":1}' timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_flags -c 1 -o gpurun_out/prof_scan_v9 python tools/stitch_sweep.py > gpurun_out/ncu_scan.log 2>&1; echo "ncu rc=$?"; ls -la gpurun_out
