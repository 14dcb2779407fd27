"""helpers shared by the CPU and GPU golden-fixture tests (tests/golden/*.json, made by make_golden.py)"""
import glob
import hashlib
import json
import os

import numpy as np

from _oracle import gen_reads, sort_recs

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.json")))


def load_case(name, tmpdir):
    """-> (fixture dict, path of the regenerated input, its bytes); the input sha256 is verified"""
    g = json.load(open(os.path.join(GOLDEN_DIR, name + ".json")))
    path = gen_reads(os.path.join(str(tmpdir), name + ".txt"), **g["case"]["gen"])
    text = open(path, "rb").read()
    assert hashlib.sha256(text).hexdigest() == g["input_sha256"], "tools/gen_reads no longer reproduces the fixture input"
    return g, path, text


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def junction_lines(recs, k):
    """'KMER d0..d4  c0..c3 csum  l0..l4 ' per record, as Junction::toString + print_kmer
    (utils/Junction.cpp:74-89, utils/Kmer.cpp:555-564) -- independent python restatement"""
    out = []
    for r in recs:
        x = int(r["kmer"])
        s = "".join("ACTG"[(x >> (2 * (k - 1 - i))) & 3] for i in range(k))
        d = " ".join(str(int(v)) for v in r["dist"])
        c = " ".join(str(int(v)) for v in r["cov"])
        l = " ".join(str(int(v)) for v in r["linked"])
        out.append(f"{s} {d}  {c} {int(sum(int(v) for v in r['cov']))}  {l} ")
    return out


def check_records(g, recs, k):
    r = sort_recs(recs[["kmer", "dist", "cov", "linked", "pad"]] if "creation_rank" in recs.dtype.names else recs)
    assert len(r) == g["n_records"]
    head = [[int(x["kmer"]), x["dist"].tolist(), x["cov"].tolist(), x["linked"].tolist()] for x in r[:1000]]
    assert head == g["records"]
    import numpy as np
    from _oracle import REC_DTYPE
    rr = np.zeros(len(r), REC_DTYPE)
    for f in ("kmer", "dist", "cov", "linked"):
        rr[f] = r[f]
    assert sha(rr) == g["records_sha256"]
    lines = sorted(junction_lines(rr, k))
    assert lines[:8] == g["junction_lines_head"]
    assert hashlib.sha256("\n".join(lines).encode()).hexdigest() == g["junctions_sorted_lines_sha256"]
