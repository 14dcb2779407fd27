"""The oracle (oracle/faucet_oracle.c) against every known answer we have for the hot path:
  * SURVEY Appendix B.1 hash / k-mer KATs (extracted from the compiled reference),
  * the reference's own scan vectors, src/newTests/ReadscanTest.cpp:102-281 (fake Bloom, j=0, spacer 8),
  * tests/golden/*.json, produced by running the unmodified reference (tests/golden/make_golden.py).
CPU only; the GPU path is held to the same fixtures in test_gpu_golden.py.
"""
import ctypes

import numpy as np
import pytest

from _golden import CASES, check_records, load_case, sha

READ = "ACGTTGCATGCCGATAGCTAGCTAGGATCGATCGTACGATCGTAGCTAGCTAGCTGATCGATCGTAGC"
KAT = [  # pos, fwd, revcomp, canonical is fwd?, h0, h1  (Bloom(1000000,31): tai = 2^20)
    (0, 0x07ad2d7236363c9c, 0x27258d8d89f4b41e, True, 0xd952a, 0x89016),
    (1, 0x1eb4b5c8d8d8f272, 0x09c96363627d2d07, False, 0xccf8d, 0xad99c),
    (2, 0x3ad2d7236363c9c9, 0x327258d8d89f4b41, False, 0xa94af, 0x1a94a),
    (3, 0x2b4b5c8d8d8f2727, 0x1c9c96363627d2d0, False, 0xf69cd, 0xeadd6),
]


def test_seeds(oracle):
    assert oracle.lib.fo_seed(0) == 0xffaa54ffe6e6e6e7
    assert oracle.lib.fo_seed(1) == 0x1140aada557088a4


@pytest.mark.parametrize("pos,fwd,rc,fwd_is_canon,h0,h1", KAT)
def test_kmer_hash_kat(oracle, pos, fwd, rc, fwd_is_canon, h0, h1):
    assert oracle.first_kmer(READ[pos:], 31) == fwd
    assert oracle.lib.fo_revcomp(fwd, 31) == rc
    c = oracle.lib.fo_canon(fwd, 31)
    assert c == (fwd if fwd_is_canon else rc) == min(fwd, rc)
    assert oracle.lib.fo_old_hash(c, 0, 20) == h0
    assert oracle.lib.fo_old_hash(c, 1, 20) == h1


def test_hash_kat_big_filter(oracle):
    assert oracle.lib.fo_old_hash(0x0123456789abcdef, 0, 33) == 0xd127affb
    assert oracle.lib.fo_old_hash(0x0123456789abcdef, 1, 33) == 0x1183769ce


def test_nt_codes(oracle):  # utils/Kmer.cpp:82-93: A0 C1 T2 G3, complement = +2 mod 4
    assert [oracle.lib.fo_nt2int(ctypes.c_char(c)) for c in b"ACTG"] == [0, 1, 2, 3]
    assert oracle.kmer_string(oracle.first_kmer("GATTACA", 7), 7) == "GATTACA"
    assert oracle.kmer_string(oracle.lib.fo_revcomp(oracle.first_kmer("GATTACA", 7), 7), 7) == "TGTAATC"


@pytest.mark.parametrize("est,sing,p1,bits,nh,lt", [  # SURVEY Appendix D (computed by the reference)
    (10**6, 10**4, 0.04125, 6, 4, 23), (4_600_000, 10**6, 0.07366, 5, 3, 25), (12_000_000, 10_000_000, 0.22943, 3, 2, 26),
    (64_000_000, 20_000_000, 0.09268, 4, 2, 28), (10**9, 2 * 10**8, 0.07046, 5, 3, 33), (3 * 10**9, 10**9, 0.09717, 4, 2, 34)])
def test_geometry_table(oracle, est, sing, p1, bits, nh, lt):
    p = oracle.lib.fo_brent_p1(est, sing, 0.04)
    assert abs(p - p1) < 1e-5
    assert oracle.geometry_optimal(est, ctypes.c_float(p).value) == (lt, nh)


def test_pair_filter_geometry(oracle):  # src/Faucet.cpp:273-274: 9 bits, 6 hashes
    assert oracle.geometry_optimal(10**6 // 20, 0.01)[1] == 6
    assert oracle.geometry_optimal(10**6 // 10, 0.01)[1] == 6


# ---- src/newTests/ReadscanTest.cpp -----------------------------------------------------------------
R1, R2, R3 = "ACGGGCGAACTTTCATAGGA", "GGCGAACTAGTCCAT", "AACTTTCATACGATT"
K1 = ["ACGGG", "CGGGC", "GGGCG", "GGCGA", "GCGAA", "CGAAC", "GAACT", "AACTT", "ACTTT", "CTTTC", "TTTCA", "TTCAT",
      "TCATA", "CATAG", "ATAGG", "TAGGA"]
VECTORS = {
    "singleReadNoJunctions": (5, [R1], K1, {"TCCTA": [0, 0, 15, 0, 1], "AACTT": [0, 0, 15, 0, 15]}),
    "singleReadOneFakeJunction": (5, [R1], K1 + ["AACTC", "ACTCC"],
                                  {"CCTAT": [0, 0, 0, 15, 3], "GAACT": [0, 0, 15, 0, 13]}),
    "LongReadNoJunctions": (5, ["ACGGGCGAACTTTCATAGGATCGCACTCAC"],
                            K1 + ["AGGAT", "GGATC", "GATCG", "ATCGC", "TCGCA", "CGCAC", "GCACT", "CACTC", "ACTCA", "CTCAC"],
                            {"TGCGA": [0, 0, 1, 0, 11], "ATCGC": [1, 0, 0, 0, 3], "CGATC": [0, 1, 0, 0, 3],
                             "GGATC": [0, 0, 0, 1, 12], "TTCAT": [12, 0, 0, 0, 15], "TTCGC": [0, 1, 0, 0, 15],
                             "GGCGA": [1, 0, 0, 0, 7]}),
    "buildFullMap": (5, [R1, R2, R3],
                     K1 + ["AACTA", "ACTAG", "CTAGT", "TAGTC", "AGTCC", "GTCCA", "TCCAT", "CATAC", "ATACG", "TACGA",
                           "ACGAT", "CGATT"],
                     {"CTAGT": [0, 8, 3, 0, 3], "TCATA": [0, 10, 0, 6, 12], "GAACT": [3, 0, 12, 0, 13]}),
    "smallDblJuncMap": (7, ["AAAAACAGCGATTC", "AAAAAGAGCGATTTA"],
                        ["AAAAACA", "AAAAAGA", "AAAACAG", "AAAAGAG", "AAACAGC", "AAAGAGC", "AACAGCG", "AAGAGCG",
                         "ACAGCGA", "AGAGCGA", "CAGCGAT", "GAGCGAT", "AGCGATT", "GCGATTT", "GCGATTC", "CGATTTA"],
                        {"AGCGATT": [0, 2, 4, 0, 1], "AATCGCT": [0, 12, 0, 12, 1]}),
}


@pytest.mark.parametrize("name", sorted(VECTORS))
def test_readscan_vectors(oracle, name):
    k, reads, kmers, expect = VECTORS[name]
    fake = sorted({oracle.lib.fo_canon(oracle.first_kmer(s, k), k) for s in kmers})
    recs, st = oracle.scan_reads(reads, k, 0, 8, fake)
    got = {oracle.kmer_string(int(r["kmer"]), k): r["dist"].tolist() for r in recs}
    assert got == expect
    assert st["n_junctions"] == len(expect)


# ---- fixtures generated from the unmodified reference ----------------------------------------------
@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_fixture(oracle, tmp_path, name):
    g, path, text = load_case(name, tmp_path)
    c = g["case"]
    p1 = ctypes.c_float(oracle.lib.fo_brent_p1(c["est"], c["sing"], 0.04)).value
    assert p1 == g["p1_float"]
    lt, nh = oracle.geometry_optimal(c["est"], p1)
    assert (lt, nh) == (g["log2_tai"], g["n_hash"])
    b1, b2, st = oracle.load_two_filters(text, c["fastq"], c["k"], lt, nh)
    assert sha(b1) == g["bloo1_sha256"] and sha(b2) == g["bloo2_sha256"]
    assert int(np.unpackbits(b2).sum()) == g["bloo2_bits"]
    sg, lg = tuple(g["spf_geom"]), tuple(g["lpf_geom"])
    assert oracle.geometry_optimal(max(1, c["est"] // 20), 0.01) == sg
    spf, lpf = np.zeros((1 << sg[0]) // 8, np.uint8), np.zeros((1 << lg[0]) // 8, np.uint8)
    recs, sst = oracle.scan(text, c["fastq"], c["paired"], c["no_cleaning"], c["k"], c["j"], c["spacer"], b2, lt, nh,
                            spf, sg, lpf, lg)
    assert sst == g["scan_stats"]
    check_records(g, recs, c["k"])
    assert sha(spf) == g["spf_sha256"] and sha(lpf) == g["lpf_sha256"]
    # the oracle's own .junctions formatter agrees with the python restatement used by check_records
    from _golden import junction_lines
    from _oracle import sort_recs
    r = sort_recs(recs)
    assert [s.rstrip("\n") for s in oracle.junction_lines(r[:50], c["k"])] == junction_lines(r[:50], c["k"])
