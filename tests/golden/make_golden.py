#!/usr/bin/env python
"""Generates tests/golden/*.json from the UNMODIFIED reference (oracle/_ref/libfaucet_ref.so, built by
oracle/Makefile from /root/reference).  Run in the build container only:  python tests/golden/make_golden.py

Each fixture pins, for one seeded synthetic input (tools/gen_reads, platform-independent RNG):
  the input's sha256, both Bloom bit arrays of pass 1 (sha256 + weight), the junction map of pass 2
  (every record, sorted by k-mer), the reference's own .junctions text (sha256 of its sorted lines),
  both pair filters (sha256) and the scan counters printed by ReadScanner::printScanSummary.
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _oracle import Ref, gen_reads, sort_recs  # noqa: E402

CASES = {
    "fq_k31_j1_paired_clean": dict(gen=dict(genome=20000, cov=20, length=100, insert=300, seed=11, err=0.005,
                                            nrate=0.002, repeats=True),
                                   fastq=1, paired=1, no_cleaning=0, k=31, j=1, spacer=100, est=20000, sing=10000),
    "fa_k25_j1_spacer40": dict(gen=dict(genome=15000, cov=15, length=150, insert=400, seed=12, err=0.01,
                                        nrate=0.004, fasta=True),
                               fastq=0, paired=0, no_cleaning=1, k=25, j=1, spacer=40, est=15000, sing=8000),
    "fq_k21_j2": dict(gen=dict(genome=12000, cov=25, length=100, insert=250, seed=13, err=0.008),
                      fastq=1, paired=1, no_cleaning=1, k=21, j=2, spacer=100, est=12000, sing=6000),
    "fq_k27_j0_errfree": dict(gen=dict(genome=30000, cov=30, length=100, insert=300, seed=14, repeats=True),
                              fastq=1, paired=1, no_cleaning=0, k=27, j=0, spacer=100, est=30000, sing=1000),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = Ref()
    for name, c in CASES.items():
        with tempfile.TemporaryDirectory() as d:
            path = gen_reads(os.path.join(d, "reads.txt"), **c["gen"])
            text = open(path, "rb").read()
            ref.set_k(c["k"])
            import ctypes
            p1 = ctypes.c_float(ref.lib.ref_brent_p1(c["est"], c["sing"], 0.04)).value
            lt, nh = ref.geometry_optimal(c["est"], p1)
            b1, b2 = ref.load_two_filters(path, c["fastq"], c["k"], lt, nh)
            sg, lg = ref.geometry_optimal(max(1, c["est"] // 20), 0.01), ref.geometry_optimal(max(1, c["est"] // 10), 0.01)
            spf = np.zeros((1 << sg[0]) // 8, np.uint8)
            lpf = np.zeros((1 << lg[0]) // 8, np.uint8)
            jpath = os.path.join(d, "out.junctions")
            recs, st = ref.scan(path, c["fastq"], c["paired"], c["no_cleaning"], c["k"], c["j"], c["spacer"], b2, lt,
                                nh, spf, sg, lpf, lg, junctions_path=jpath)
            lines = sorted(open(jpath).read().splitlines())
            r = sort_recs(recs)
            out = {
                "case": {k: v for k, v in c.items()},
                "input_sha256": hashlib.sha256(text).hexdigest(), "input_bytes": len(text),
                "p1_float": p1, "log2_tai": lt, "n_hash": nh, "spf_geom": list(sg), "lpf_geom": list(lg),
                "bloo1_sha256": sha(b1), "bloo2_sha256": sha(b2),
                "bloo1_bits": int(np.unpackbits(b1).sum()), "bloo2_bits": int(np.unpackbits(b2).sum()),
                "scan_stats": st,
                "junctions_sorted_lines_sha256": hashlib.sha256("\n".join(lines).encode()).hexdigest(),
                "junction_lines_head": lines[:8],
                "records_sha256": sha(r),
                "n_records": len(r),
                # every record when the map is small, else the first 1000 (records_sha256 pins the rest)
                "records": [[int(x["kmer"]), x["dist"].tolist(), x["cov"].tolist(), x["linked"].tolist()]
                            for x in r[:1000]],
                "spf_sha256": sha(spf), "lpf_sha256": sha(lpf),
                "spf_bits": int(np.unpackbits(spf).sum()), "lpf_bits": int(np.unpackbits(lpf).sum()),
            }
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(out, f, separators=(",", ":"))
        print(name, "junctions", len(recs), "bloo2 bits", out["bloo2_bits"], "spf/lpf bits", out["spf_bits"], out["lpf_bits"])


if __name__ == "__main__":
    main()
