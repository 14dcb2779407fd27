#!/usr/bin/env python
"""GPU experiment (not part of the product): times the stitch of a bench workload under different knobs.
usage: SWEEP='{"epoch0":4096};{"epoch_mode":0}' python tools/stitch_sweep.py [workload]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import bench
import faucet_b200 as fb

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
k = w["k"]
_, lt, nh = fb.geometry_from_reads(w["est"], w["sing"], bench.FP)
path = bench.gen_dataset(w, seed=1)
raw = np.fromfile(path, dtype=np.uint8)
n = min(raw.size, int(os.environ.get("SWEEP_BYTES", 2 << 30)))
n = int(np.flatnonzero(raw[:n] == 10)[-1]) + 1 if n < raw.size else n
n -= 0
dev = torch.from_numpy(raw[:n]).cuda()
CONFIGS = [dict(), dict(epoch_mode=1)]
if os.environ.get("SWEEP"):
    CONFIGS = [json.loads(x) for x in os.environ["SWEEP"].split(";")]
DEFAULT = dict(stitch_exec=1, epoch_recheck=1, flow_chunk=1 << 22, epoch_mode=0, epoch0=8192, epoch_max=1 << 20, epoch_switch_pct=30, epoch_shrink_pct=14, epoch_grow_pct=6,
               stitch_blocks=3, stitch_shrink_den=4, stitch_grow_den=10, stitch_w_max=1 << 15, res_log2=24, table_cap0=1 << 22)
ref = None
for cfg in CONFIGS:
    for kk, v in {**DEFAULT, **cfg}.items():
        fb.set_tuning(kk, v)
    s = fb.Session(k, lt, nh, j=1, max_spacer_dist=100, max_text_bytes=n)
    s.set_text((dev.data_ptr(), n), device=True)
    s.parse(True); s.load(); s.get_bloom(to_host=False); s.scan_flags()
    s.sync()
    ts = []
    for _ in range(3):
        s.set_profiling(True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nj = s.stitch(True, True)
        s.sync()
        wall = (time.perf_counter() - t0) * 1e3
        ts.append((wall, s.kernel_ms("stitch")[0], s.kernel_ms("stitch_dry")[0], s.kernel_ms("stitch_verify")[0], s.kernel_ms("stitch_flow_prep")[0]))
        s.set_profiling(False)
    recs, st = s.junctions()
    t = fb.timings()
    sig = (nj, st["nb_processed"], st["nb_skipped"], st["nb_jcheck_kmer"], int(recs["kmer"].sum() & 0xffffffff), int(recs["cov"].astype(np.int64).sum()))
    if ref is None:
        ref = sig
    best = min(ts)
    print(json.dumps({"cfg": cfg, "wall_ms": round(best[0], 2), "ordered_ms": round(best[1], 2), "dry_ms": round(best[2], 2), "verify_ms": round(best[3], 2), "flow_prep_ms": round(best[4], 2),
                      "rounds": t["stitch_rounds"], "deferred": t["stitch_deferred"],
                      "epochs": [t["epochs_exact"], t["epochs_classify"]], "exact_records": t["exact_records"], "dry_records": t["dry_records"],
                      "iterations": t["epoch_iterations"], "nonquiet": t["nonquiet_records"], "writers": t["writer_records"],
                      "fallbacks": t["epoch_fallbacks"], "phase": t["stitch_phase_ns"], "reads": st["reads_processed"], "same_result": sig == ref}), flush=True)
    s.close()
