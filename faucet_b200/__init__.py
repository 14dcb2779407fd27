"""faucet_b200: B200-native implementation of Faucet's two-pass streaming k-mer hot path.

The product is libfaucet_gpu.so (hand-written sm_100a CUDA behind the C ABI in include/faucet_gpu.h)
plus the C++11 host adaptors under faucet_b200/host/.  This Python package is only the ctypes binding
used by tests/ and bench.py; it holds no compute and has no CPU fallback.
"""
from ._lib import (FaucetError, JunctionRec, LoadStats, ScanStats, Session, device_count, geometry_2_hash,
                   geometry_from_reads, geometry_optimal, lib, load_two_filters, load_two_filters_mem, plan_shards, query_ext_masks, scan,
                   scan_mem, scan_retained, set_batch_bytes, set_epoch_limit, set_tuning, timings, REC_DTYPE)

__all__ = ["FaucetError", "JunctionRec", "LoadStats", "ScanStats", "Session", "device_count", "geometry_2_hash",
           "geometry_from_reads", "geometry_optimal", "lib", "load_two_filters", "load_two_filters_mem", "plan_shards", "query_ext_masks", "scan",
           "scan_mem", "scan_retained", "set_batch_bytes", "set_epoch_limit", "set_tuning", "timings", "REC_DTYPE"]
