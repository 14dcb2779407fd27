"""The C++ host driver (faucet_b200/bin/faucet) against the UNMODIFIED reference binary (oracle/_ref/faucet,
built from /root/reference by oracle/Makefile; it travels to the GPU box prebuilt): same command line,
byte-identical <prefix>.bloom, <prefix>.junctions and pair-filter files."""
import filecmp
import os
import subprocess

import pytest

from _oracle import REF_BIN, gen_reads

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "faucet_b200", "bin", "faucet")


def _run(exe, reads, prefix, extra, timeout=600):
    cmd = [exe, "-read_load_file", reads, "-read_scan_file", reads, "-size_kmer", "31", "-max_read_length", "100",
           "-estimated_kmers", "80000", "-singletons", "40000", "-file_prefix", prefix] + extra
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("extra", [["--fastq", "--paired_ends", "--no_cleaning"], ["--fastq", "--paired_ends", "--two_hash"],
                                   ["--fastq", "-j", "2", "-max_spacer_dist", "30", "-fp", "0.02", "--no_cleaning"]])
def test_cli_files_match_reference_binary(tmp_path, extra):
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/faucet was not built")
    assert os.path.exists(CLI), "faucet_b200/bin/faucet missing: run __graft_entry__.build()"
    reads = gen_reads(str(tmp_path / "r.fq"), genome=80000, cov=25, length=100, insert=300, seed=51, err=0.005, nrate=0.001,
                      repeats=True)
    ours = _run(CLI, reads, str(tmp_path / "ours"), extra)
    assert ours.returncode == 0, ours.stderr + ours.stdout[-2000:]
    ref = _run(REF_BIN, reads, str(tmp_path / "ref"), extra)  # may die later, in graph cleaning: files come first
    suffixes = [".bloom", ".junctions"] + ([] if "--no_cleaning" in extra else [".short_pair_filter", ".long_pair_filter"])
    for suf in suffixes:
        a, b = str(tmp_path / "ours") + suf, str(tmp_path / "ref") + suf
        assert os.path.exists(b), f"reference wrote no {suf}: {ref.stdout[-500:]} {ref.stderr[-500:]}"
        assert os.path.getsize(a) > 0 and filecmp.cmp(a, b, shallow=False), f"{suf} differs from the reference's file"
    # the counters the reference prints
    for key in ("Reads processed:", "Unambiguous reads:", "Number of junctions:"):
        mine = [l for l in ours.stdout.splitlines() if l.startswith(key)]
        theirs = [l for l in ref.stdout.splitlines() if l.startswith(key)]
        assert mine and mine == theirs[:len(mine)], (key, mine, theirs)


def test_cli_restart_from_bloom_file(tmp_path):
    """-bloom_file: geometry re-derived from -fp (and --two_hash honoured), as getBloomFilterFromFile does"""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/faucet was not built")
    reads = gen_reads(str(tmp_path / "r.fq"), genome=50000, cov=20, length=100, insert=300, seed=52, err=0.004)
    # a filter whose geometry is the same under both derivations: produce it with the reference itself
    base = ["--fastq", "--no_cleaning", "-fp", "0.05"]
    ref1 = _run(REF_BIN, reads, str(tmp_path / "ref"), base + ["--just_load_bloom"])
    assert os.path.exists(str(tmp_path / "ref.bloom")), ref1.stdout[-500:]
    ours = _run(CLI, reads, str(tmp_path / "ours"), base + ["-bloom_file", str(tmp_path / "ref.bloom")])
    ref2 = _run(REF_BIN, reads, str(tmp_path / "ref2"), base + ["-bloom_file", str(tmp_path / "ref.bloom")])
    assert ours.returncode == 0, ours.stderr
    assert filecmp.cmp(str(tmp_path / "ours.junctions"), str(tmp_path / "ref2.junctions"), shallow=False)


DROPIN = os.path.join(ROOT, "faucet_b200", "bin", "faucet_dropin")
ALL_FILES = [".bloom", ".junctions", ".short_pair_filter", ".long_pair_filter", ".cleaned_contigs.fasta",
             ".cleaned_graph_unitigs.fastg"]


def _run2(exe, reads, prefix, k, max_len, est, sing, extra, timeout=900):
    cmd = [exe, "-read_load_file", reads, "-read_scan_file", reads, "-size_kmer", str(k), "-max_read_length", str(max_len),
           "-estimated_kmers", str(est), "-singletons", str(sing), "-file_prefix", prefix] + extra
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)


def _fastg_text(path):
    """the reference names FASTG nodes after heap addresses (NODE_0x55d0..., src/ContigGraph.cpp:1470-1500), which change
    from run to run of the SAME binary: number them by first appearance before comparing"""
    import re
    ids = {}
    return re.sub(r"NODE_0x[0-9a-f]+", lambda m: "NODE_%d" % ids.setdefault(m.group(0), len(ids)), open(path).read())


def _same_files(tmp_path, suffixes, ref):
    for suf in suffixes:
        a, b = str(tmp_path / "ours") + suf, str(tmp_path / "ref") + suf
        assert os.path.exists(b), f"reference wrote no {suf}: {ref.stdout[-500:]} {ref.stderr[-500:]}"
        assert os.path.exists(a), f"no {suf} from the GPU path"
        assert os.path.getsize(a) > 0
        if suf.endswith(".fastg"):
            assert _fastg_text(a) == _fastg_text(b), f"{suf} differs from the reference's file (node addresses normalised)"
        else:
            assert filecmp.cmp(a, b, shallow=False), f"{suf} differs from the reference's file"


@pytest.mark.parametrize("extra,suffixes", [
    (["--fastq", "--paired_ends"], ALL_FILES),                       # cleaning on: pair filters, contigs, unitig graph
    (["--fastq", "--paired_ends", "--no_cleaning"], ALL_FILES[:2]),
    (["--fastq", "-j", "0", "-max_spacer_dist", "40"], [s for s in ALL_FILES if s != ".long_pair_filter"]),
])
def test_dropin_binary_reproduces_reference_contigs(tmp_path, extra, suffixes):
    """faucet_dropin = the reference's OWN main() and graph stage (compiled unmodified) linked against
    faucet_b200/host/dropin/adaptors.cpp, which runs load_two_filters and ReadScanner::scanReads on the GPU.
    Planted repeats keep the reference's cleaning alive.  Every file the reference writes -- down to the
    cleaned contigs -- must be byte-identical."""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/faucet was not built")
    assert os.path.exists(DROPIN), "faucet_b200/bin/faucet_dropin missing: run __graft_entry__.build() where /root/reference is mounted"
    reads = gen_reads(str(tmp_path / "r.fq"), genome=200000, cov=30, length=100, insert=300, seed=61, err=0.005, repeats=True)
    ours = _run2(DROPIN, reads, str(tmp_path / "ours"), 31, 100, 2400000, 2000000, extra)
    assert ours.returncode == 0, ours.stderr[-2000:] + ours.stdout[-2000:]
    ref = _run2(REF_BIN, reads, str(tmp_path / "ref"), 31, 100, 2400000, 2000000, extra)
    assert ref.returncode == 0, ref.stderr[-2000:]
    _same_files(tmp_path, suffixes, ref)
    if "--no_cleaning" not in extra:
        assert ">" in open(str(tmp_path / "ours.cleaned_contigs.fasta")).read()


@pytest.mark.parametrize("variant", ["e0", "e5"])
def test_config1_full_size_matches_reference_binary(tmp_path, variant):
    """BASELINE configs[0] at full size (1 Mbp, 30x, 100 bp interleaved paired-end FASTQ, k = 31: 21 M k-mers per
    pass), error-free and with 0.5 % errors + planted repeats: every file byte-identical to the reference binary's"""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/faucet was not built")
    if variant == "e0":  # a repeat-free genome: the reference's cleaning dies, so stop after the graph build
        kw, est, sing, extra, sufs, exe = {}, 1000000, 10000, ["--fastq", "--paired_ends", "--no_cleaning"], ALL_FILES[:2], CLI
    else:
        kw, est, sing, extra, sufs, exe = dict(err=0.005, repeats=True), 12000000, 10000000, ["--fastq", "--paired_ends"], ALL_FILES, DROPIN
    reads = gen_reads(str(tmp_path / "c1.fq"), genome=1000000, cov=30, length=100, insert=300, seed=1, **kw)
    ours = _run2(exe, reads, str(tmp_path / "ours"), 31, 100, est, sing, extra)
    assert ours.returncode == 0, ours.stderr[-2000:] + ours.stdout[-2000:]
    ref = _run2(REF_BIN, reads, str(tmp_path / "ref"), 31, 100, est, sing, extra)
    _same_files(tmp_path, sufs, ref)
    for key in ("Reads processed:", "Unambiguous reads:", "Number of junctions:"):
        mine = [l for l in ours.stdout.splitlines() if l.startswith(key)]
        theirs = [l for l in ref.stdout.splitlines() if l.startswith(key)]
        assert mine and mine == theirs[:len(mine)], (key, mine, theirs)
