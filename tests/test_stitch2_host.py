"""The thread-per-record junction walk of the CUDA stitch (faucet_b200/csrc/stitch2_walk.cuh) is
__host__ __device__ code: tests/stitch2_host.cpp compiles it for the CPU around a host environment
and this file holds it to the oracle -- junction records in creation order, scan counters and both
pair filters, bit for bit -- so the walk logic is checked without a GPU.  (The GPU kernel that runs
the same code under the reservation schedule is covered by tests/test_gpu_parity.py.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from _oracle import REC_DTYPE, ScanStats, gen_reads

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def s2h(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("s2h") / "libstitch2_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "stitch2_host.cpp")])
    lib = C.CDLL(so)
    lib.s2h_scan.restype = C.c_int
    lib.s2h_free.argtypes = [C.c_void_p]
    return lib


def _run(lib, text, fastq, paired, no_cleaning, k, j, spacer, bloo2, lt, nh, spf=None, spf_geom=(0, 0), lpf=None,
         lpf_geom=(0, 0)):
    recs = C.c_void_p()
    n = C.c_uint64()
    st = ScanStats()
    u8 = C.POINTER(C.c_uint8)
    p = lambda a: None if a is None else a.ctypes.data_as(u8)
    rc = lib.s2h_scan(C.c_char_p(text), C.c_size_t(len(text)), int(fastq), int(paired), int(no_cleaning), k, j, spacer,
                      p(bloo2), lt, nh, p(spf), spf_geom[0], spf_geom[1], p(lpf), lpf_geom[0], lpf_geom[1],
                      C.byref(recs), C.byref(n), C.byref(st), None)
    assert rc == 0
    arr = np.zeros(n.value, REC_DTYPE)
    if n.value:
        C.memmove(arr.ctypes.data, recs, n.value * 24)
    lib.s2h_free(recs)
    return arr, st.as_dict()


def _same(a, b):
    for f in ("kmer", "dist", "cov", "linked"):
        assert np.array_equal(a[f], b[f]), f


CASES = [
    # (name, gen_reads kwargs, fastq, k, j, spacer)
    ("fq_k31_j1", dict(genome=60000, cov=30, length=100, insert=300, seed=3, err=0.005, nrate=0.002, repeats=True), True, 31, 1, 100),
    ("fq_k21_j2", dict(genome=30000, cov=25, length=100, insert=250, seed=4, err=0.01, nrate=0.003, repeats=True), True, 21, 2, 100),
    ("fa_k25_j0_spacer", dict(genome=40000, cov=20, length=150, insert=400, seed=5, err=0.01, nrate=0.004, fasta=True), False, 25, 0, 20),
    ("fq_k31_150_spacers_fire", dict(genome=50000, cov=40, length=150, insert=350, seed=6, err=0.002), True, 31, 1, 100),
    ("fq_k32_lower", dict(genome=20000, cov=20, length=120, insert=300, seed=8, err=0.01, lower=True), True, 32, 1, 30),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_walk_matches_oracle(s2h, oracle, tmp_path, case):
    name, kw, fastq, k, j, spacer = case
    text = open(gen_reads(str(tmp_path / name), **kw), "rb").read()
    lt, nh = oracle.geometry_optimal(kw["genome"] * 2, 0.04)
    _, b2, _ = oracle.load_two_filters(text, fastq, k, lt, nh)
    orecs, ostats = oracle.scan(text, fastq, True, 1, k, j, spacer, b2, lt, nh)
    hrecs, hstats = _run(s2h, text, fastq, True, 1, k, j, spacer, b2, lt, nh)
    assert hstats == ostats
    _same(hrecs, orecs)


def test_walk_pair_filters_match_oracle(s2h, oracle, tmp_path):
    kw = dict(genome=50000, cov=30, length=100, insert=300, seed=12, err=0.005, nrate=0.001, repeats=True)
    text = open(gen_reads(str(tmp_path / "pf.fq"), **kw), "rb").read()
    k, j = 31, 1
    lt, nh = oracle.geometry_optimal(100000, 0.04)
    _, b2, _ = oracle.load_two_filters(text, True, k, lt, nh)
    geom_s, geom_l = oracle.geometry_optimal(5000, 0.01), oracle.geometry_optimal(10000, 0.01)
    o_spf, o_lpf = np.zeros((1 << geom_s[0]) // 8, np.uint8), np.zeros((1 << geom_l[0]) // 8, np.uint8)
    h_spf, h_lpf = o_spf.copy(), o_lpf.copy()
    orecs, ostats = oracle.scan(text, True, True, 0, k, j, 100, b2, lt, nh, o_spf, geom_s, o_lpf, geom_l)
    hrecs, hstats = _run(s2h, text, True, True, 0, k, j, 100, b2, lt, nh, h_spf, geom_s, h_lpf, geom_l)
    assert hstats == ostats
    _same(hrecs, orecs)
    assert o_spf.any() and o_lpf.any()
    assert np.array_equal(h_spf, o_spf) and np.array_equal(h_lpf, o_lpf)
