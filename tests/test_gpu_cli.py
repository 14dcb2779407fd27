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
