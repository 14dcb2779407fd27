#!/usr/bin/env python
"""GPU experiment (not part of the product): times the stitch kernel of the C2 workload under
different window / occupancy knobs.  usage: python tools/stitch_sweep.py [workload]"""
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
dev = torch.from_numpy(raw).cuda()
n = raw.size
CONFIGS = [dict(), dict(stitch_blocks=3), dict(stitch_shrink_den=2, stitch_grow_den=4), dict(stitch_w_max=1024)]
if os.environ.get("SWEEP"):
    CONFIGS = [json.loads(x) for x in os.environ["SWEEP"].split(";")]
DEFAULT = dict(stitch_impl=1, stitch_blocks=3, stitch_shrink_den=4, stitch_grow_den=10, stitch_w_max=1 << 15, res_log2=24,
               table_cap0=1 << 22)
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
        nj = s.stitch(True, True)
        s.sync()
        ts.append(s.kernel_ms("stitch")[0])
        s.set_profiling(False)
    recs, st = s.junctions()
    t = fb.timings()
    sig = (nj, st["nb_processed"], st["nb_skipped"], int(recs["kmer"].sum() & 0xffffffff))
    if ref is None:
        ref = sig
    print(json.dumps({"cfg": cfg, "stitch_ms": min(ts), "rounds": t["stitch_rounds"], "deferred": t["stitch_deferred"],
                      "phase_us_per_round": [round(x / 1e3 / max(1, t["stitch_rounds"]), 2) for x in t["stitch_phase_ns"]],
                      "same_result": sig == ref}), flush=True)
    s.close()
