"""GPU parity: libfaucet_gpu.so (through its C ABI) against the oracle on the same seeded inputs.

Bit-exact bar: identical bloo1/bloo2 byte arrays, identical junction records (key, dist, cov, linked)
in identical creation order, identical pair filters and identical scan counters.
"""
import ctypes
import os

import numpy as np
import pytest

from _oracle import gen_reads, sort_recs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fb():
    import faucet_b200
    if faucet_b200.device_count() == 0:
        pytest.fail("no CUDA device: the gpu-marked tests need the B200 box")
    return faucet_b200


def _dataset(tmp_path_factory, name, **kw):
    p = str(tmp_path_factory.mktemp("reads") / name)
    gen_reads(p, **kw)
    with open(p, "rb") as f:
        return p, f.read()


@pytest.fixture(scope="module")
def small_fq(tmp_path_factory):
    # 100 kbp genome, 30x, 0.5% errors, N bases (multi-segment reads), planted repeats
    return _dataset(tmp_path_factory, "small.fq", genome=100000, cov=30, length=100, insert=300, seed=3,
                    err=0.005, nrate=0.002, repeats=True)


@pytest.fixture(scope="module")
def small_fa(tmp_path_factory):
    return _dataset(tmp_path_factory, "small.fa", genome=60000, cov=20, length=150, insert=400, seed=5,
                    err=0.01, nrate=0.004, fasta=True)


EPOCH_DEFAULTS = {"epoch_mode": 0, "epoch0": 8192, "epoch_max": 1 << 20, "epoch_recheck": 1, "stitch_exec": 1, "flow_chunk": 1 << 22,
                  "dry_lazy": 2}
EPOCH_SCHEDULES = {
    "adaptive": {"epoch_mode": 1},                                      # ordered epochs first, then classify epochs
    "ordered_only": {},                                                 # the default: every record through the dataflow executor
    "ordered_small_chunks": {"flow_chunk": 333},                        # ... with a dependency sort every 333 records
    "rounds_only": {"epoch_mode": 0, "stitch_exec": 0},                 # ... of the round-based ordered kernel
    "classify_rounds": {"epoch_mode": 2, "epoch0": 300, "epoch_max": 5000, "stitch_exec": 0},
    "flow_small_chunks": {"epoch_mode": 1, "epoch0": 1024, "flow_chunk": 700},  # dependency sort every 700 records
    "classify_tiny": {"epoch_mode": 2, "epoch0": 192, "epoch_max": 3000},  # classify / execute / verify / apply from record 0 on
    "classify_join": {"epoch_mode": 2, "epoch0": 500, "epoch_max": 8000, "epoch_recheck": 0},  # tainted records all join the exact set
    "classify_parked": {"epoch_mode": 2, "epoch0": 400, "epoch_max": 6000, "dry_lazy": 0},  # read-only walks on parked lookups (big tables)
}


@pytest.fixture(params=sorted(EPOCH_SCHEDULES))
def impl(request, fb):
    """every epoch schedule of the stitch (faucet_b200/csrc/stitch.cuh) is held to the same bar"""
    for name, v in EPOCH_SCHEDULES[request.param].items():
        fb.set_tuning(name, v)
    yield request.param
    for name, v in EPOCH_DEFAULTS.items():
        fb.set_tuning(name, v)


def _geom(oracle, est, sing, fp=0.04):
    p1 = ctypes.c_float(oracle.lib.fo_brent_p1(est, sing, fp)).value
    return oracle.geometry_optimal(est, p1)


def _strip(recs):
    """drop creation_rank/pad so GPU (32-byte) and oracle (24-byte) records compare field by field"""
    return [(int(r["kmer"]), bytes(r["dist"]), bytes(r["cov"]), bytes(r["linked"])) for r in recs]


@pytest.mark.parametrize("k", [31, 21, 32])
def test_load_matches_oracle(fb, oracle, small_fq, k):
    _, text = small_fq
    lt, nh = _geom(oracle, 100000, 50000)
    o1, o2, ost = oracle.load_two_filters(text, True, k, lt, nh)
    g2, g1, gst = fb.load_two_filters_mem(text, True, k, lt, nh, want_bloo1=True)
    assert np.array_equal(g1, o1)
    assert np.array_equal(g2, o2)
    assert gst.kmers == ost.kmers
    assert gst.unambiguous_reads == ost.unambiguous_reads
    assert gst.reads_processed == ost.reads_processed
    assert gst.weight1 == ost.weight1 and gst.weight2 == ost.weight2


@pytest.mark.parametrize("nh,lt", [(1, 18), (2, 19), (3, 20), (4, 17), (6, 21), (7, 22), (10, 22)])
def test_load_hash_counts(fb, oracle, small_fa, nh, lt):
    _, text = small_fa
    o1, o2, _ = oracle.load_two_filters(text, False, 27, lt, nh)
    g2, g1, _ = fb.load_two_filters_mem(text, False, 27, lt, nh, want_bloo1=True)
    assert np.array_equal(g1, o1) and np.array_equal(g2, o2)


def test_load_multibatch_and_epochs(fb, oracle, small_fq):
    """tiny batches (cuts at record boundaries) and a tiny stamp epoch give the same filters"""
    _, text = small_fq
    lt, nh = _geom(oracle, 100000, 50000)
    o1, o2, ost = oracle.load_two_filters(text, True, 31, lt, nh)
    try:
        fb.set_batch_bytes(200_000)
        fb.set_epoch_limit(500_000)
        g2, g1, gst = fb.load_two_filters_mem(text, True, 31, lt, nh, want_bloo1=True)
    finally:
        fb.set_batch_bytes(256 << 20)
        fb.set_epoch_limit(0xfffffffe)
    assert np.array_equal(g1, o1) and np.array_equal(g2, o2)
    assert gst.kmers == ost.kmers and gst.reads_processed == ost.reads_processed


@pytest.mark.parametrize("j", [0, 1, 2])
@pytest.mark.parametrize("no_cleaning", [1, 0])
def test_scan_matches_oracle(fb, oracle, small_fq, j, no_cleaning, impl):
    _, text = small_fq
    k = 31
    lt, nh = _geom(oracle, 100000, 50000)
    _, b2, _ = oracle.load_two_filters(text, True, k, lt, nh)
    sg, lg = oracle.geometry_optimal(100000 // 20, 0.01), oracle.geometry_optimal(100000 // 10, 0.01)
    ospf, olpf = np.zeros((1 << sg[0]) // 8, np.uint8), np.zeros((1 << lg[0]) // 8, np.uint8)
    gspf, glpf = ospf.copy(), olpf.copy()
    orecs, ost = oracle.scan(text, True, True, no_cleaning, k, j, 100, b2, lt, nh, ospf, sg, olpf, lg)
    grecs, gst = fb.scan_mem(text, True, True, no_cleaning, k, j, 100, b2, lt, nh, gspf, sg, glpf, lg)
    assert gst == ost
    assert _strip(grecs) == _strip(orecs)          # same records in the same creation order
    assert list(grecs["creation_rank"]) == list(range(len(grecs)))
    assert np.array_equal(gspf, ospf) and np.array_equal(glpf, olpf)


def test_scan_fasta_spacers_multibatch(fb, oracle, small_fa, impl):
    """150 bp reads with max_spacer_dist 40 (spacers fire), FASTA, unpaired, tiny batches"""
    _, text = small_fa
    k = 25
    lt, nh = 21, 3
    _, b2, _ = oracle.load_two_filters(text, False, k, lt, nh)
    orecs, ost = oracle.scan(text, False, False, 1, k, 1, 40, b2, lt, nh)
    try:
        fb.set_batch_bytes(150_000)
        grecs, gst = fb.scan_mem(text, False, False, 1, k, 1, 40, b2, lt, nh)
    finally:
        fb.set_batch_bytes(256 << 20)
    assert gst == ost
    assert _strip(grecs) == _strip(orecs)


@pytest.mark.parametrize("tail", [b"", b"\n", b"@last", b"@lastACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT",
                                  b"@h\nACGTNACG", b"@h\nACGTACGTACGTAGCTAGCTAGCTAGCATCGATCGATCAGCTAGC\n+", b"\n\n"])
def test_ragged_tails(fb, oracle, small_fq, tail, impl):
    """truncated / unterminated inputs, including the header-reused-as-sequence getline quirk"""
    _, text = small_fq
    text = text[:40_000]
    text = text[:text.rfind(b"\n@") + 1] + tail
    k, lt, nh = 15, 18, 3
    o1, o2, ost = oracle.load_two_filters(text, True, k, lt, nh)
    g2, g1, gst = fb.load_two_filters_mem(text, True, k, lt, nh, want_bloo1=True)
    assert np.array_equal(g1, o1) and np.array_equal(g2, o2)
    assert gst.reads_processed == ost.reads_processed and gst.kmers == ost.kmers
    orecs, osst = oracle.scan(text, True, True, 1, k, 1, 100, o2, lt, nh)
    grecs, gsst = fb.scan_mem(text, True, True, 1, k, 1, 100, o2, lt, nh)
    assert gsst == osst and _strip(grecs) == _strip(orecs)


def test_empty_and_tiny_inputs(fb, oracle, impl):
    k, lt, nh = 31, 16, 4
    for text in (b"", b"\n", b">x\n", b">x\nACGT\n", b">x\n" + b"ACGT" * 8 + b"\n",
                 b">x\n" + b"ACGT" * 7 + b"ACG" + b"\n", b">x\r\n" + b"ACGT" * 10 + b"\r\n"):
        o1, o2, ost = oracle.load_two_filters(text, False, k, lt, nh)
        g2, g1, gst = fb.load_two_filters_mem(text, False, k, lt, nh, want_bloo1=True)
        assert np.array_equal(g1, o1) and np.array_equal(g2, o2), text
        assert gst.kmers == ost.kmers and gst.reads_processed == ost.reads_processed, text
        orecs, osst = oracle.scan(text, False, False, 1, k, 1, 100, o2, lt, nh)
        grecs, gsst = fb.scan_mem(text, False, False, 1, k, 1, 100, o2, lt, nh)
        assert gsst == osst and _strip(grecs) == _strip(orecs), text


def test_session_stage_api(fb, oracle, small_fq):
    """the device-resident stage API gives the same answer as the whole-pass calls"""
    _, text = small_fq
    k = 31
    lt, nh = _geom(oracle, 100000, 50000)
    _, o2, _ = oracle.load_two_filters(text, True, k, lt, nh)
    orecs, ost = oracle.scan(text, True, True, 1, k, 1, 100, o2, lt, nh)
    s = fb.Session(k, lt, nh, j=1, max_spacer_dist=100, max_text_bytes=len(text))
    s.set_text(text)
    s.parse(True)
    s.load()
    b2, _ = s.get_bloom()
    assert np.array_equal(b2, o2)
    s.scan_flags()
    assert s.stitch(True, True) == len(orecs)
    grecs, gst = s.junctions()
    assert _strip(grecs) == _strip(orecs)
    assert s.launches > 0
    s.close()


@pytest.mark.parametrize("knobs", [
    {"table_cap0": 64},                                   # junction table grows by rehash many times
    {"ext_cap0": 64},                                     # extension-list buffer drained / regrown mid-batch
    {"res_log2": 8, "stitch_w0": 4096},                   # reservation collisions: most records get deferred
    {"stitch_w0": 1, "stitch_w_max": 1},                  # one record per round == plain sequential order
    {"stitch_w0": 32768, "stitch_w_max": 32768},          # window far larger than the genome supports
    {"table_cap0": 256, "ext_cap0": 256, "res_log2": 10, "stitch_w0": 512},
    {"epoch_mode": 2, "epoch0": 64, "epoch_max": 64},     # classify epochs of 64 records from the empty table on
    {"epoch_mode": 2, "epoch0": 1 << 20},                 # ONE classify epoch per batch: the exact set closes over nearly everything
    {"epoch_mode": 2, "epoch0": 1000, "table_cap0": 64},  # the table has to grow inside classify epochs (fallback to the ordered kernel)
    {"epoch_mode": 2, "epoch0": 700, "ext_cap0": 64},     # extension lists drained inside the exact runs and between apply launches
    {"epoch_mode": 2, "epoch0": 500, "res_log2": 8},      # every write taints most of the epoch (256 slots)
    {"epoch_mode": 1, "epoch0": 256, "epoch_max": 4096, "epoch_switch_pct": 100},  # adaptive, switching to classify at once
    {"stitch_w0": 1 << 17, "stitch_w_max": 1 << 17},      # more window than the warp-per-record grid has warps
])
def test_stitch_schedule_is_invisible(fb, oracle, small_fq, knobs, impl):
    """whatever the round schedule of the GPU stitch (window, collisions, growth, drains), the junction
    map, counters and pair filters equal the sequential oracle's"""
    _, text = small_fq
    k, j = 31, 1
    lt, nh = _geom(oracle, 100000, 50000)
    _, b2, _ = oracle.load_two_filters(text, True, k, lt, nh)
    sg, lg = oracle.geometry_optimal(100000 // 20, 0.01), oracle.geometry_optimal(100000 // 10, 0.01)
    ospf, olpf = np.zeros((1 << sg[0]) // 8, np.uint8), np.zeros((1 << lg[0]) // 8, np.uint8)
    gspf, glpf = ospf.copy(), olpf.copy()
    orecs, ost = oracle.scan(text, True, True, 0, k, j, 100, b2, lt, nh, ospf, sg, olpf, lg)
    defaults = {"table_cap0": 1 << 22, "ext_cap0": 1 << 24, "res_log2": 24, "stitch_w0": 2048, "stitch_w_max": 1 << 15,
                "epoch_switch_pct": 30, **EPOCH_DEFAULTS}
    try:
        for name, v in knobs.items():
            fb.set_tuning(name, v)
        fb.set_batch_bytes(700_000)  # several batches: the table and the mate carry live across them
        grecs, gst = fb.scan_mem(text, True, True, 0, k, j, 100, b2, lt, nh, gspf, sg, glpf, lg)
        tim = fb.timings()
    finally:
        for name, v in defaults.items():
            fb.set_tuning(name, v)
        fb.set_batch_bytes(256 << 20)
    assert gst == ost
    assert _strip(grecs) == _strip(orecs)
    assert np.array_equal(gspf, ospf) and np.array_equal(glpf, olpf)
    assert tim["exact_records"] > 0


def test_stitch_repetitive_reads(fb, oracle, tmp_path_factory, impl):
    """every read drawn from a 2 kbp genome at 400x: nearly all records conflict with their neighbours"""
    p, text = _dataset(tmp_path_factory, "rep.fq", genome=2000, cov=400, length=100, insert=300, seed=17, err=0.01)
    k, lt, nh = 21, 18, 3
    _, b2, _ = oracle.load_two_filters(text, True, k, lt, nh)
    for j in (0, 1):
        orecs, ost = oracle.scan(text, True, True, 1, k, j, 100, b2, lt, nh)
        grecs, gst = fb.scan_mem(text, True, True, 1, k, j, 100, b2, lt, nh)
        assert gst == ost and _strip(grecs) == _strip(orecs)


def test_stitch_mixed_line_lengths(fb, oracle, tmp_path_factory, impl):
    """100 bp and 300 bp reads of one genome interleaved file-wise: the long lines (280 k-mer positions) take the
    direct (not shared-memory staged) path of the ordered kernel"""
    _, short = _dataset(tmp_path_factory, "mix_s.fq", genome=30000, cov=15, length=100, insert=300, seed=21, err=0.005, nrate=0.002)
    _, long_ = _dataset(tmp_path_factory, "mix_l.fq", genome=30000, cov=15, length=300, insert=700, seed=21, err=0.005, nrate=0.002)
    a, b = short.split(b"\n")[:-1], long_.split(b"\n")[:-1]
    recs = []
    for i in range(0, max(len(a), len(b)), 8):  # two mate pairs of one file, then two of the other
        recs += a[i:i + 8] + b[i:i + 8]
    text = b"\n".join(recs) + b"\n"
    k, j, lt, nh = 21, 1, 20, 3
    o1, o2, _ = oracle.load_two_filters(text, True, k, lt, nh)
    g2, _, _ = fb.load_two_filters_mem(text, True, k, lt, nh)
    assert np.array_equal(g2, o2)
    sg, lg = oracle.geometry_optimal(3000, 0.01), oracle.geometry_optimal(6000, 0.01)
    for no_cleaning in (1, 0):
        ospf, olpf = np.zeros((1 << sg[0]) // 8, np.uint8), np.zeros((1 << lg[0]) // 8, np.uint8)
        gspf, glpf = ospf.copy(), olpf.copy()
        orecs, ost = oracle.scan(text, True, True, no_cleaning, k, j, 60, o2, lt, nh, ospf, sg, olpf, lg)
        grecs, gst = fb.scan_mem(text, True, True, no_cleaning, k, j, 60, o2, lt, nh, gspf, sg, glpf, lg)
        assert gst == ost and _strip(grecs) == _strip(orecs)
        assert np.array_equal(gspf, ospf) and np.array_equal(glpf, olpf)


def test_scan_k32_all_g_key(fb, oracle, impl):
    """k = 32: the all-'G' k-mer has the bit pattern of the table's empty marker"""
    reads = [b"G" * 70, b"ACGT" * 5 + b"G" * 40 + b"TTGCA" * 4, b"C" * 70, b"G" * 50 + b"A" + b"G" * 40]
    text = b"".join(b">r%d\n%s\n" % (i, r) for i, r in enumerate(reads * 3))
    k, lt, nh = 32, 16, 3
    _, b2, _ = oracle.load_two_filters(text, False, k, lt, nh)
    g2, _, _ = fb.load_two_filters_mem(text, False, k, lt, nh)
    assert np.array_equal(g2, b2)
    for j in (0, 1):
        orecs, ost = oracle.scan(text, False, False, 1, k, j, 100, b2, lt, nh)
        grecs, gst = fb.scan_mem(text, False, False, 1, k, j, 100, b2, lt, nh)
        assert gst == ost and _strip(grecs) == _strip(orecs)


@pytest.mark.parametrize("sub0,sub", [(32, 32 * 64), (4096, 4096), (1 << 20, 64 << 20)])
def test_load_sub_batches(fb, oracle, small_fq, sub0, sub):
    """the load pass may cut a batch into sub-batches of any size (even one 32-byte word) without changing a bit"""
    _, text = small_fq
    text = text[:text.find(b"\n@", 400_000) + 1]
    lt, nh = _geom(oracle, 100000, 50000)
    o1, o2, ost = oracle.load_two_filters(text, True, 31, lt, nh)
    try:
        fb.set_tuning("load_sub_bytes0", sub0)
        fb.set_tuning("load_sub_bytes", sub)
        g2, g1, gst = fb.load_two_filters_mem(text, True, 31, lt, nh, want_bloo1=True)
    finally:
        fb.set_tuning("load_sub_bytes0", 1 << 20)
        fb.set_tuning("load_sub_bytes", 64 << 20)
    assert np.array_equal(g1, o1) and np.array_equal(g2, o2)
    assert gst.kmers == ost.kmers and gst.unambiguous_reads == ost.unambiguous_reads


def test_scan_retained_matches_scan_mem(fb, oracle, small_fq, impl):
    """pass 2 on the planes pass 1 left in HBM (no text, no second parse) == pass 2 on the text == the oracle;
    several batches, both pair filters"""
    _, text = small_fq
    k, j = 31, 1
    lt, nh = _geom(oracle, 100000, 50000)
    sg, lg = oracle.geometry_optimal(100000 // 20, 0.01), oracle.geometry_optimal(100000 // 10, 0.01)
    ospf, olpf = np.zeros((1 << sg[0]) // 8, np.uint8), np.zeros((1 << lg[0]) // 8, np.uint8)
    gspf, glpf = ospf.copy(), olpf.copy()
    _, o2, _ = oracle.load_two_filters(text, True, k, lt, nh)
    orecs, ost = oracle.scan(text, True, True, 0, k, j, 100, o2, lt, nh, ospf, sg, olpf, lg)
    with pytest.raises(fb.FaucetError):  # nothing retained yet
        fb.scan_retained(True, 0, k, j, 100, None, lt, nh)
    try:
        fb.set_tuning("retain_planes", 1)
        fb.set_batch_bytes(900_000)
        g2, _, _ = fb.load_two_filters_mem(text, True, k, lt, nh)
        assert np.array_equal(g2, o2)
        grecs, gst = fb.scan_retained(True, 0, k, j, 100, None, lt, nh, gspf, sg, glpf, lg)  # device copy of bloo2
        assert gst == ost and _strip(grecs) == _strip(orecs)
        assert np.array_equal(gspf, ospf) and np.array_equal(glpf, olpf)
        grecs, gst = fb.scan_retained(True, 1, k, j, 100, g2, lt, nh)  # again, bloo2 from the host this time
        orecs1, ost1 = oracle.scan(text, True, True, 1, k, j, 100, o2, lt, nh)
        assert gst == ost1 and _strip(grecs) == _strip(orecs1)
        fb.set_tuning("retain_budget", 1000)  # too small: pass 1 still works, pass 2 must ask for the text
        fb.load_two_filters_mem(text, True, k, lt, nh)
        with pytest.raises(fb.FaucetError):
            fb.scan_retained(True, 1, k, j, 100, None, lt, nh)
    finally:
        fb.set_tuning("retain_planes", 0)
        fb.set_tuning("retain_budget", 64 << 30)
        fb.set_batch_bytes(256 << 20)


def test_path_entry_points_stream_the_file(fb, oracle, small_fq, tmp_path):
    """faucet_gpu_load_two_filters / faucet_gpu_scan read the file through a reader thread and pinned staging
    buffers, chunk by chunk: many tiny chunks, a ragged last record and a missing file (= zero reads, as the
    reference's unchecked ifstream, utils/Bloom.cpp:268-269) all match the oracle on the same bytes"""
    p, text = small_fq
    ragged = str(tmp_path / "ragged.fq")
    rtext = text[:text.rfind(b"\n@") + 1] + b"@tail\nACGTACGTACGTAGCTAGCTAGCTAGCATCGATCGATCAGCTAGC"
    open(ragged, "wb").write(rtext)
    k, j = 31, 1
    lt, nh = _geom(oracle, 100000, 50000)
    try:
        fb.set_batch_bytes(300_000)
        for path, t in ((p, text), (ragged, rtext), (str(tmp_path / "missing.fq"), b"")):
            o1, o2, ost = oracle.load_two_filters(t, True, k, lt, nh)
            g2, g1, gst = fb.load_two_filters(path, True, k, lt, nh, want_bloo1=True)
            assert np.array_equal(g1, o1) and np.array_equal(g2, o2)
            assert gst.kmers == ost.kmers and gst.reads_processed == ost.reads_processed
            orecs, osst = oracle.scan(t, True, True, 1, k, j, 100, o2, lt, nh)
            grecs, gsst = fb.scan(path, True, True, 1, k, j, 100, g2, lt, nh)
            assert gsst == osst and _strip(grecs) == _strip(orecs)
    finally:
        fb.set_batch_bytes(256 << 20)


@pytest.mark.parametrize("j", [0, 1, 2])
def test_scan_flags_memo_is_invisible(fb, oracle, small_fq, j):
    """scan_flags looks up the per-k-mer extension masks it has already computed (scan_memo, the default) or
    computes every position from the Bloom filter (scan_memo = 0): same junction map, same counters; and a memo
    that is far too small for the input (it is a cache) changes nothing either"""
    _, text = small_fq
    k = 31
    lt, nh = _geom(oracle, 100000, 50000)
    _, b2, _ = oracle.load_two_filters(text, True, k, lt, nh)
    orecs, ost = oracle.scan(text, True, True, 1, k, j, 100, b2, lt, nh)
    try:
        for memo in (0, 1):
            fb.set_tuning("scan_memo", memo)
            fb.set_batch_bytes(500_000)  # several batches share one memo
            grecs, gst = fb.scan_mem(text, True, True, 1, k, j, 100, b2, lt, nh)
            assert gst == ost and _strip(grecs) == _strip(orecs), memo
    finally:
        fb.set_tuning("scan_memo", 1)
        fb.set_batch_bytes(256 << 20)


def test_load_saturated_kmer_cache_is_invisible(fb, oracle, small_fq):
    """pass 1 with its cache of saturated k-mers forced on (it is meant for filters that do not fit L2): same bloo1 and
    bloo2, same counters, in one batch and in many; and pass 2 right after it re-uses the buffer for its own cache"""
    _, text = small_fq
    lt, nh = _geom(oracle, 100000, 50000)
    o1, o2, ost = oracle.load_two_filters(text, True, 31, lt, nh)
    orecs, osst = oracle.scan(text, True, True, 1, 31, 1, 100, o2, lt, nh)
    try:
        fb.set_tuning("load_memo_log2", 0)
        for batch in (256 << 20, 300_000):
            fb.set_batch_bytes(batch)
            g2, g1, gst = fb.load_two_filters_mem(text, True, 31, lt, nh, want_bloo1=True)
            assert np.array_equal(g1, o1) and np.array_equal(g2, o2)
            assert gst.kmers == ost.kmers and gst.weight1 == ost.weight1 and gst.weight2 == ost.weight2
            grecs, gsst = fb.scan_mem(text, True, True, 1, 31, 1, 100, g2, lt, nh)
            assert gsst == osst and _strip(grecs) == _strip(orecs)
    finally:
        fb.set_tuning("load_memo_log2", 29)
        fb.set_batch_bytes(256 << 20)


def test_scan_retained_j_change_clears_the_memo(fb, oracle, small_fq):
    """the memo caches depth-j masks: scan_retained with bloo2 = None and another j must not reuse them"""
    _, text = small_fq
    k = 31
    lt, nh = _geom(oracle, 100000, 50000)
    _, o2, _ = oracle.load_two_filters(text, True, k, lt, nh)
    try:
        fb.set_tuning("retain_planes", 1)
        fb.load_two_filters_mem(text, True, k, lt, nh)
        for j in (1, 0, 2, 1):
            orecs, ost = oracle.scan(text, True, True, 1, k, j, 100, o2, lt, nh)
            grecs, gst = fb.scan_retained(True, 1, k, j, 100, None, lt, nh)
            assert gst == ost and _strip(grecs) == _strip(orecs), j
    finally:
        fb.set_tuning("retain_planes", 0)


@pytest.mark.parametrize("j", [3, 4])
def test_scan_deep_jcheck(fb, oracle, tmp_path_factory, j):
    """j = 3 and 4, the deepest the reference's JChecker scratch arrays allow (utils/JChecker.cpp:93-94)"""
    _, text = _dataset(tmp_path_factory, "deep.fq", genome=20000, cov=25, length=100, insert=300, seed=31, err=0.01)
    k, lt, nh = 25, 17, 2  # a full filter: many false-positive branches to chase
    _, b2, _ = oracle.load_two_filters(text, True, k, lt, nh)
    for memo in (1, 0):
        try:
            fb.set_tuning("scan_memo", memo)
            orecs, ost = oracle.scan(text, True, True, 1, k, j, 100, b2, lt, nh)
            grecs, gst = fb.scan_mem(text, True, True, 1, k, j, 100, b2, lt, nh)
        finally:
            fb.set_tuning("scan_memo", 1)
        assert gst == ost and _strip(grecs) == _strip(orecs), memo


def test_lower_case_bases_split_reads(fb, oracle, tmp_path_factory, impl):
    """lower-case bases are not valid nucleotides (utils/Kmer.cpp:50-60): they split a read like an N does"""
    _, text = _dataset(tmp_path_factory, "lower.fq", genome=40000, cov=25, length=120, insert=300, seed=41, err=0.005,
                       nrate=0.004, lower=True)
    assert any(c in text for c in (b"a", b"c", b"g", b"t"))
    k, lt, nh = 27, 19, 3
    o1, o2, ost = oracle.load_two_filters(text, True, k, lt, nh)
    g2, g1, gst = fb.load_two_filters_mem(text, True, k, lt, nh, want_bloo1=True)
    assert np.array_equal(g1, o1) and np.array_equal(g2, o2)
    assert gst.kmers == ost.kmers and gst.unambiguous_reads == ost.unambiguous_reads
    for no_cleaning in (1, 0):
        sg, lg = oracle.geometry_optimal(2000, 0.01), oracle.geometry_optimal(4000, 0.01)
        ospf, olpf = np.zeros((1 << sg[0]) // 8, np.uint8), np.zeros((1 << lg[0]) // 8, np.uint8)
        gspf, glpf = ospf.copy(), olpf.copy()
        orecs, osst = oracle.scan(text, True, True, no_cleaning, k, 1, 100, o2, lt, nh, ospf, sg, olpf, lg)
        grecs, gsst = fb.scan_mem(text, True, True, no_cleaning, k, 1, 100, o2, lt, nh, gspf, sg, glpf, lg)
        assert gsst == osst and _strip(grecs) == _strip(orecs)
        assert np.array_equal(gspf, ospf) and np.array_equal(glpf, olpf)


def test_epochs_report_their_work(fb, oracle, tmp_path_factory):
    """a deep-coverage stream spends most records in classify epochs; the exact set stays a small share of them"""
    _, text = _dataset(tmp_path_factory, "deep_cov.fq", genome=60000, cov=120, length=150, insert=400, seed=51)
    k = 31
    lt, nh = _geom(oracle, 60000, 20000)
    _, b2, _ = oracle.load_two_filters(text, True, k, lt, nh)
    orecs, ost = oracle.scan(text, True, True, 1, k, 1, 100, b2, lt, nh)
    try:
        fb.set_tuning("epoch_mode", 1)
        fb.set_tuning("epoch0", 1024)
        grecs, gst = fb.scan_mem(text, True, True, 1, k, 1, 100, b2, lt, nh)
        tim = fb.timings()
    finally:
        fb.set_tuning("epoch0", 8192)
        fb.set_tuning("epoch_mode", 0)
    assert gst == ost and _strip(grecs) == _strip(orecs)
    assert tim["epochs_classify"] > 0 and tim["dry_records"] > ost["reads_processed"] // 2
    assert tim["exact_records"] < ost["reads_processed"]


@pytest.mark.parametrize("j", [0, 1, 2])
def test_query_ext_masks_matches_getValidJExtension(fb, oracle, ref, small_fq, j):
    """faucet_gpu_query_ext_masks (batched Bloom walks for the contig build) against the reference's own
    JunctionMap::getValidJExtension (utils/JunctionMap.cpp:474-490) on k-mers of the reads, both orientations, plus
    random k-mers (mostly non-members)"""
    _, text = small_fq
    k = 31
    lt, nh = _geom(oracle, 100000, 50000)
    _, b2, _ = oracle.load_two_filters(text, True, k, lt, nh)
    seqs = [l for l in text.split(b"\n")[1:4000:4] if b"N" not in l and len(l) >= k + 1]
    kmers = []
    for s in seqs[:300]:
        for p in range(0, len(s) - k, 7):
            x = oracle.first_kmer(s[p:p + k].decode(), k)
            kmers += [x, oracle.lib.fo_revcomp(x, k)]
    rng = np.random.default_rng(9)
    kmers += [int(v) for v in rng.integers(0, 1 << 62, size=2000, dtype=np.uint64)]
    kmers = np.array(kmers, np.uint64)
    masks = fb.query_ext_masks(kmers, k, j, b2, lt, nh)
    valid = masks >> 4
    ours = np.where(valid == 0, -1, np.where((valid & (valid - 1)) != 0, -2, np.log2(np.maximum(valid, 1)).astype(np.int32)))
    theirs = ref.valid_j_extension(kmers, k, j, b2, lt, nh)
    assert np.array_equal(ours, theirs)
    assert ((masks & 15) | valid == (masks & 15)).all()  # a valid extension is a member
    # the device copy the call left behind serves the next one (bloo2 = None)
    assert np.array_equal(fb.query_ext_masks(kmers[:100], k, j, None, lt, nh), masks[:100])
