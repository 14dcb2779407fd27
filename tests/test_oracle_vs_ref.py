"""Pins the oracle (oracle/faucet_oracle.c) to the UNMODIFIED reference compiled into oracle/_ref
(oracle/Makefile + oracle/ref_shim.cpp): k-mer codec, hash, geometry, both Bloom arrays of pass 1, the
junction map, scan counters and pair filters of pass 2.  Skipped when oracle/_ref was not built (the
reference sources only exist in the build container; tests/golden/*.json carry its answers elsewhere).
"""
import ctypes
import random

import numpy as np
import pytest

from _oracle import gen_reads, sort_recs


def _p1(lib_fn, est, sing, fp=0.04):
    return ctypes.c_float(lib_fn(est, sing, fp)).value


@pytest.mark.parametrize("k", [5, 17, 27, 31, 32])
def test_codec_and_hash(oracle, ref, k):
    ref.set_k(k)
    rnd = random.Random(k)
    mask = (1 << (2 * k)) - 1
    for _ in range(300):
        x = rnd.getrandbits(64) & mask
        assert oracle.lib.fo_revcomp(x, k) == ref.lib.ref_revcomp(x)
        assert oracle.lib.fo_canon(x, k) == ref.lib.ref_get_canon(x)
        for lt in (10, 20, 33):
            for i in (0, 1):
                assert oracle.lib.fo_old_hash(x, i, lt) == ref.lib.ref_old_hash(lt, x, i)
    for i in range(10):
        assert oracle.lib.fo_seed(i) == ref.lib.ref_seed(i)


@pytest.mark.parametrize("est,sing", [(10**6, 10**4), (4_600_000, 10**6), (12_000_000, 10_000_000), (64_000_000, 20_000_000),
                                      (10**9, 2 * 10**8), (3 * 10**9, 10**9), (1000, 10), (35, 3), (5 * 10**5, 5 * 10**5)])
def test_geometry(oracle, ref, est, sing):
    ref.set_k(31)
    for fp in (0.04, 0.01, 0.1):
        po, pr = oracle.lib.fo_brent_p1(est, sing, fp), ref.lib.ref_brent_p1(est, sing, fp)
        assert po == pr
        f = ctypes.c_float(po).value
        assert oracle.geometry_optimal(est, f) == ref.geometry_optimal(est, f)
        assert oracle.geometry_2_hash(est, f) == ref.geometry_2_hash(est, f)
    assert oracle.geometry_optimal(max(1, est // 20), 0.01) == ref.geometry_optimal(max(1, est // 20), 0.01)


CASES = [
    dict(gen=dict(genome=40000, cov=25, length=100, insert=300, seed=21, err=0.005, nrate=0.002, repeats=True),
         fastq=1, paired=1, k=31, j=1, spacer=100),
    dict(gen=dict(genome=30000, cov=20, length=150, insert=400, seed=22, err=0.01, nrate=0.004, fasta=True),
         fastq=0, paired=0, k=25, j=1, spacer=40),
    dict(gen=dict(genome=20000, cov=20, length=100, insert=250, seed=23, err=0.01, nrate=0.003, lower=True),
         fastq=1, paired=1, k=21, j=2, spacer=100),
    dict(gen=dict(genome=20000, cov=30, length=120, insert=300, seed=24), fastq=1, paired=1, k=32, j=0, spacer=30),
]


@pytest.mark.parametrize("ci", range(len(CASES)))
@pytest.mark.parametrize("no_cleaning", [1, 0])
def test_load_and_scan(oracle, ref, tmp_path, ci, no_cleaning):
    c = CASES[ci]
    path = gen_reads(str(tmp_path / "r.txt"), **c["gen"])
    text = open(path, "rb").read()
    est = c["gen"]["genome"]
    lt, nh = oracle.geometry_optimal(est, _p1(oracle.lib.fo_brent_p1, est, est // 2))
    r1, r2 = ref.load_two_filters(path, c["fastq"], c["k"], lt, nh)
    o1, o2, _ = oracle.load_two_filters(text, c["fastq"], c["k"], lt, nh)
    assert np.array_equal(o1, r1) and np.array_equal(o2, r2)
    sg, lg = oracle.geometry_optimal(max(1, est // 20), 0.01), oracle.geometry_optimal(max(1, est // 10), 0.01)
    ospf, olpf = np.zeros((1 << sg[0]) // 8, np.uint8), np.zeros((1 << lg[0]) // 8, np.uint8)
    rspf, rlpf = ospf.copy(), olpf.copy()
    rrecs, rst = ref.scan(path, c["fastq"], c["paired"], no_cleaning, c["k"], c["j"], c["spacer"], r2, lt, nh, rspf, sg, rlpf, lg)
    orecs, ost = oracle.scan(text, c["fastq"], c["paired"], no_cleaning, c["k"], c["j"], c["spacer"], o2, lt, nh, ospf, sg, olpf, lg)
    assert ost == rst
    assert np.array_equal(sort_recs(orecs), sort_recs(rrecs))
    assert np.array_equal(ospf, rspf) and np.array_equal(olpf, rlpf)


def test_creation_order_rebuilds_reference_iteration_order(oracle, ref, tmp_path):
    """inserting the ORACLE's records into a libstdc++ unordered_map in the oracle's CREATION order must reproduce the
    reference's .junctions line order: that is how the host adaptors rebuild the map from the GPU's records"""
    c = CASES[0]
    path = gen_reads(str(tmp_path / "r.txt"), **c["gen"])
    text = open(path, "rb").read()
    lt, nh = 19, 3
    _, b2 = ref.load_two_filters(path, 1, c["k"], lt, nh)
    jp = str(tmp_path / "ref.junctions")
    ref.scan(path, 1, 1, 1, c["k"], 1, 100, b2, lt, nh, junctions_path=jp)
    orecs, _ = oracle.scan(text, 1, 1, 1, c["k"], 1, 100, b2, lt, nh)
    ref_lines = [l.split(" ")[0] for l in open(jp).read().splitlines()]
    rebuilt = ref.iteration_order(orecs["kmer"])  # the oracle returns its records in creation order
    assert [oracle.kmer_string(int(x), c["k"]) for x in rebuilt] == ref_lines
    # and the order matters: inserting the same keys sorted gives another file order
    assert [oracle.kmer_string(int(x), c["k"]) for x in ref.iteration_order(np.sort(orecs["kmer"]))] != ref_lines


def test_fake_bloom_vectors_agree(oracle, ref):
    """the ReadscanTest reads through the reference's fakify() path and through the oracle's fake-set path"""
    from test_golden import VECTORS
    for name, (k, reads, kmers, expect) in VECTORS.items():
        ref.set_k(k)
        fake = sorted({oracle.lib.fo_canon(oracle.first_kmer(s, k), k) for s in kmers})
        rrecs, rst = ref.scan_fake(reads, k, 0, 8, fake)
        orecs, ost = oracle.scan_reads(reads, k, 0, 8, fake)
        assert np.array_equal(sort_recs(orecs), sort_recs(rrecs)), name
        for f in ("nb_jcheck_kmer", "nb_no_juncs", "nb_processed", "nb_skipped", "reads_no_errors", "unambiguous_reads"):
            assert ost[f] == rst[f], (name, f)


def test_fake_junction_that_grows_the_map(oracle, ref, tmp_path):
    """regression: the junction that fills the oracle's record array (the 1025th here) is a mid-read FAKE
    junction; the oracle once took the record pointer before the creation reallocated the array and lost
    that junction's coverage / distances.  Found by tests/test_stitch2_host.py; pinned here."""
    kw = dict(genome=60000, cov=30, length=100, insert=300, seed=3, err=0.005, nrate=0.002, repeats=True)
    full = open(gen_reads(str(tmp_path / "full.fq"), **kw), "rb").read()
    text = b"\n".join(full.split(b"\n")[:4 * 2568]) + b"\n"
    path = str(tmp_path / "trunc.fq")
    open(path, "wb").write(text)
    k, j = 31, 1
    lt, nh = oracle.geometry_optimal(120000, 0.04)
    _, b2, _ = oracle.load_two_filters(full, True, k, lt, nh)
    rrecs, rst = ref.scan(path, True, True, 1, k, j, 100, b2, lt, nh)
    orecs, ost = oracle.scan(text, True, True, 1, k, j, 100, b2, lt, nh)
    assert len(orecs) == 1025 and ost == rst
    assert np.array_equal(sort_recs(orecs), sort_recs(rrecs))
    assert orecs[-1]["cov"].sum() == 1  # the fake junction kept its coverage
