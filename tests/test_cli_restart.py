"""-junctions_file restart of the host CLI (faucet_b200/bin/faucet): JunctionMap::buildFromFile
(/root/reference/utils/JunctionMap.cpp:619-639) + Bloom::load of the pair filters.  Both streaming passes are
skipped, so this runs without a GPU."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "faucet_b200", "bin", "faucet")


def test_junctions_file_restart_reads_the_reference_format(tmp_path, oracle):
    if not os.path.exists(CLI):
        pytest.skip("faucet_b200/bin/faucet not built")
    k = 21
    rng = np.random.default_rng(5)
    kmers = sorted(set(int(x) for x in rng.integers(0, 1 << (2 * k), size=300)))
    lines = []
    for i, km in enumerate(kmers):
        s = "".join("ACTG"[(km >> (2 * (k - 1 - p))) & 3] for p in range(k))
        d, c, l = [(i + q) % 200 for q in range(5)], [(i * 3 + q) % 256 for q in range(4)], [(i >> q) & 1 for q in range(5)]
        lines.append(f"{s} {' '.join(map(str, d))}  {' '.join(map(str, c))} {sum(c)}  {' '.join(map(str, l))} \n")
    lines.append(lines[0])  # a repeated k-mer replaces the earlier line (std::unordered_map operator[])
    prefix = str(tmp_path / "run")
    open(prefix + ".junctions", "w").writelines(lines)
    spf = np.zeros(1 << 10, np.uint8); spf[::7] = 0xff
    spf.tofile(prefix + ".short_pair_filter")
    bloom = str(tmp_path / "b.bloom")
    np.zeros(1 << 12, np.uint8).tofile(bloom)
    out = subprocess.run([CLI, "-read_load_file", "none", "-read_scan_file", "none", "-size_kmer", str(k), "-max_read_length", "100",
                          "-estimated_kmers", "4000", "-singletons", "100", "-file_prefix", str(tmp_path / "o"), "-bloom_file", bloom,
                          "-junctions_file", prefix], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    assert f"Number of junctions: {len(kmers)}" in out.stdout
    assert "Weight of short pair filter:" in out.stdout
