"""Full-size checks (BASELINE configs[1]: 4.6 Mbp genome, 100x, 150 bp paired-end FASTQ, k = 31; 967 MB of
text, 368 M k-mers per pass) through the C ABI.  The oracle needs minutes per pass at this size, so the
bar here is made of size-independent properties of the reference's algorithm, plus the oracle itself on a
prefix of the same file:

  * batching is invisible: 256 MiB and 48 MiB device batches give byte-identical Bloom filters and
    junction records (creation order included);
  * load(T || T).bloo2 == load(T).bloo1 | load(T).bloo2: on the second copy every k-mer is already in
    bloo1, so load_two_filters (utils/Bloom.cpp:288-299) adds every k-mer to bloo2;
  * every junction key is a member of bloo2 (a junction is only ever created on a k-mer of a valid
    sub-read, src/ReadScanner.cpp:233-257), coverage never exceeds its u8 saturation, dist fits the
    half-step range of a 150 bp read, creation ranks are dense;
  * the scan counters add up: every half-step a cursor moves over is either processed or skipped.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fb():
    import faucet_b200
    if faucet_b200.device_count() == 0:
        pytest.fail("no CUDA device: the gpu-marked tests need the B200 box")
    return faucet_b200


@pytest.fixture(scope="module")
def c2():
    import bench
    w = bench.WORKLOADS["c2"]
    path = bench.gen_dataset(w, seed=1)
    return w, np.fromfile(path, dtype=np.uint8)


def _digest(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_c2_full_size_properties(fb, oracle, c2):
    import bench
    w, raw = c2
    k = w["k"]
    _, lt, nh = fb.geometry_from_reads(w["est"], w["sing"], bench.FP)
    text = raw.tobytes()
    n_reads = text.count(b"\n") // 4
    # ---- pass 1 at two batch sizes
    b2, b1, st = fb.load_two_filters_mem(text, True, k, lt, nh, want_bloo1=True)
    assert st.reads_processed == n_reads and st.kmers == n_reads * (w["length"] - k + 1)
    try:
        fb.set_batch_bytes(48 << 20)
        b2s, b1s, sts = fb.load_two_filters_mem(text, True, k, lt, nh, want_bloo1=True)
        assert _digest(b2s) == _digest(b2) and _digest(b1s) == _digest(b1) and sts.kmers == st.kmers
        # ---- the doubled stream
        d2, d1, dst = fb.load_two_filters_mem(text + text, True, k, lt, nh, want_bloo1=True)
    finally:
        fb.set_batch_bytes(256 << 20)
    assert dst.kmers == 2 * st.kmers
    assert np.array_equal(d2, b1 | b2), "second copy of the stream must put every k-mer into bloo2"
    assert np.array_equal(d1, b1), "bloo1 only grows on k-mers it does not contain"
    # ---- pass 2 at two batch sizes
    recs, sst = fb.scan_mem(text, True, True, True, k, bench.J, bench.MAX_SPACER, b2, lt, nh)
    try:
        fb.set_batch_bytes(48 << 20)
        recs_s, sst_s = fb.scan_mem(text, True, True, True, k, bench.J, bench.MAX_SPACER, b2, lt, nh)
    finally:
        fb.set_batch_bytes(256 << 20)
    assert sst_s == sst
    for f in ("kmer", "dist", "cov", "linked", "creation_rank"):
        assert np.array_equal(recs_s[f], recs[f]), f
    assert sst["reads_processed"] == n_reads and sst["n_junctions"] == len(recs) > 0
    assert list(recs["creation_rank"][:5]) == [0, 1, 2, 3, 4] and int(recs["creation_rank"][-1]) == len(recs) - 1
    assert len(np.unique(recs["kmer"])) == len(recs)
    # every junction key is a Bloom member (oracle's hash; a 20 k sample keeps this in seconds)
    mask = (1 << lt) - 1
    lib = oracle.lib
    lib.fo_old_hash.restype = C.c_uint64
    lib.fo_canon.restype = C.c_uint64
    rng = np.random.default_rng(5)
    for i in rng.choice(len(recs), size=min(20000, len(recs)), replace=False):
        c = lib.fo_canon(C.c_uint64(int(recs["kmer"][i])), k)
        h0, h1 = lib.fo_old_hash(C.c_uint64(c), 0, lt), lib.fo_old_hash(C.c_uint64(c), 1, lt)
        for t in range(nh):
            p = (h0 + t * h1) & mask
            assert b2[p >> 3] & (1 << (p & 7)), f"junction {i} is not in bloo2"
    max_half_steps = 2 * (w["length"] - k) + 1
    assert int(recs["dist"].max()) <= max_half_steps
    # a cursor only ever stands on the half-steps 2j+1 .. of its sub-read: processed + skipped is bounded by them
    assert sst["nb_processed"] + sst["nb_skipped"] <= sst["reads_no_errors"] * (max_half_steps + 255)
    assert sst["nb_processed"] >= sst["reads_no_errors"] - sst["nb_no_juncs"]


def test_c2_prefix_matches_oracle(fb, oracle, c2):
    """the first 40 k reads of the very file the bench runs on, bit for bit against the oracle (full-size Bloom geometry)"""
    import bench
    w, raw = c2
    k = w["k"]
    _, lt, nh = fb.geometry_from_reads(w["est"], w["sing"], bench.FP)
    text = raw.tobytes()
    cut = 0
    for _ in range(4 * 40000):
        cut = text.index(b"\n", cut) + 1
    text = text[:cut]
    o1, o2, ost = oracle.load_two_filters(text, True, k, lt, nh)
    g2, g1, gst = fb.load_two_filters_mem(text, True, k, lt, nh, want_bloo1=True)
    assert np.array_equal(g1, o1) and np.array_equal(g2, o2) and gst.kmers == ost.kmers
    orecs, osst = oracle.scan(text, True, True, 1, k, bench.J, bench.MAX_SPACER, o2, lt, nh)
    grecs, gsst = fb.scan_mem(text, True, True, 1, k, bench.J, bench.MAX_SPACER, g2, lt, nh)
    assert gsst == osst
    for f in ("kmer", "dist", "cov", "linked"):
        assert np.array_equal(grecs[f], orecs[f]), f


@pytest.mark.parametrize("log2_tai,n_hash", [(33, 3), (34, 2)])
def test_big_filter_prefix_matches_oracle(fb, oracle, log2_tai, n_hash):
    """BASELINE configs[3] / [4] geometry: 2^33- and 2^34-bit filters (1 and 2 GiB, bit positions beyond 32 bits, the
    saturated-k-mer cache of pass 1 on by default from 2^29 bits) on the first 40 k reads of a 100 bp stream, both
    filters and the junction map bit for bit against the oracle"""
    import bench
    from _oracle import gen_reads
    d = os.environ.get("FAUCET_BENCH_TMP", "/tmp/faucet_bench")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, "bigfilter_prefix.fq")
    if not os.path.exists(path):
        gen_reads(path, genome=2_000_000, cov=2, length=100, insert=300, seed=4, err=0.002)
    text = open(path, "rb").read()
    k = 31
    o1, o2, ost = oracle.load_two_filters(text, True, k, log2_tai, n_hash)
    g2, g1, gst = fb.load_two_filters_mem(text, True, k, log2_tai, n_hash, want_bloo1=True)
    assert gst.kmers == ost.kmers
    assert np.array_equal(g2, o2) and np.array_equal(g1, o1)
    del g1, o1
    orecs, osst = oracle.scan(text, True, True, 1, k, bench.J, bench.MAX_SPACER, o2, log2_tai, n_hash)
    grecs, gsst = fb.scan_mem(text, True, True, 1, k, bench.J, bench.MAX_SPACER, g2, log2_tai, n_hash)
    assert gsst == osst
    for f in ("kmer", "dist", "cov", "linked"):
        assert np.array_equal(grecs[f], orecs[f]), f
