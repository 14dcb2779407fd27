set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
SWEEP=${SWEEP:-'{"stitch_impl":1};{"stitch_impl":2};{"stitch_impl":2,"stitch_shrink_den":2,"stitch_grow_den":4,"stitch_w_max":65536};{"stitch_impl":2,"stitch_shrink_den":2,"stitch_grow_den":3,"stitch_w_max":16384};{"stitch_impl":2,"stitch_shrink_den":3,"stitch_grow_den":6,"stitch_w_max":65536};{"stitch_impl":2,"stitch_shrink_den":3,"stitch_grow_den":6,"stitch_w_max":8192}'} timeout 600 python tools/stitch_sweep.py > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"; tail -8 gpurun_out/sweep.log
