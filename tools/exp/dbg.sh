FAUCET_SCAN_DEBUG=1 SWEEP='{"stitch_impl":1}' timeout 600 python tools/stitch_sweep.py 2>&1 | grep "scan_flags debug" | head -3
