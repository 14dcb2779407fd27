"""GPU path (through the C ABI) against the fixtures produced by the unmodified reference
(tests/golden/*.json): Bloom arrays, junction records, .junctions text, pair filters, counters."""
import numpy as np
import pytest

from _golden import CASES, check_records, load_case, sha

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASES)
def test_gpu_matches_reference_fixture(tmp_path, name):
    import faucet_b200 as fb
    assert fb.device_count() > 0, "gpu-marked test without a CUDA device"
    g, path, text = load_case(name, tmp_path)
    c = g["case"]
    p1, lt, nh = fb.geometry_from_reads(c["est"], c["sing"], 0.04)
    assert (lt, nh) == (g["log2_tai"], g["n_hash"])
    b2, b1, st = fb.load_two_filters(path, c["fastq"], c["k"], lt, nh, want_bloo1=True)  # file-path entry point
    assert sha(b1) == g["bloo1_sha256"] and sha(b2) == g["bloo2_sha256"]
    sg, lg = tuple(g["spf_geom"]), tuple(g["lpf_geom"])
    spf, lpf = np.zeros((1 << sg[0]) // 8, np.uint8), np.zeros((1 << lg[0]) // 8, np.uint8)
    recs, sst = fb.scan(path, c["fastq"], c["paired"], c["no_cleaning"], c["k"], c["j"], c["spacer"], b2, lt, nh, spf,
                        sg, lpf, lg)
    assert sst == g["scan_stats"]
    check_records(g, recs, c["k"])
    assert sha(spf) == g["spf_sha256"] and sha(lpf) == g["lpf_sha256"]
