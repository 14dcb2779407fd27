"""The drop-in boundary without a GPU: libfaucet_gpu.so loads, exports every symbol include/faucet_gpu.h
declares, its host-side geometry code matches the oracle, and compute entry points FAIL LOUDLY when no
CUDA device is visible (there is no CPU fallback in the product)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = ""
    inc = os.path.join(ROOT, "include")
    for f in sorted(os.listdir(inc)):
        if f.endswith(".h"):
            src += open(os.path.join(inc, f)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(faucet_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import faucet_b200 as fb
    names = _declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(fb.lib, n)]
    assert not missing, missing


def test_product_never_touches_the_oracle():
    """nothing under faucet_b200/ or include/ may reference oracle/ (the oracle is test infrastructure)"""
    bad = []
    for base in ("faucet_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                    s = open(os.path.join(d, f), errors="replace").read()
                    if re.search(r"oracle/|liboracle|faucet_oracle|_ref/", s):
                        bad.append(os.path.join(d, f))
    assert not bad, bad


@pytest.mark.parametrize("est,sing", [(10**6, 10**4), (4_600_000, 10**6), (64_000_000, 20_000_000), (10**9, 2 * 10**8),
                                      (3 * 10**9, 10**9), (50000, 20000), (35, 3)])
def test_geometry_matches_oracle(oracle, est, sing):
    import faucet_b200 as fb
    for fp in (0.04, 0.01, 0.2):
        p1, lt, nh = fb.geometry_from_reads(est, sing, fp)
        assert p1 == oracle.lib.fo_brent_p1(est, sing, fp)
        assert (lt, nh) == oracle.geometry_optimal(est, ctypes.c_float(p1).value)
        assert fb.geometry_optimal(est, fp) == oracle.geometry_optimal(est, fp)
        assert fb.geometry_2_hash(est, fp) == oracle.geometry_2_hash(est, fp)


def test_no_device_is_an_error_not_a_fallback():
    import faucet_b200 as fb
    if fb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(fb.FaucetError, match="no CUDA device"):
        fb.load_two_filters_mem(b">x\nACGTACGTACGTACGTACGTACGTACGTACGTACGT\n", False, 31, 16, 4)
    with pytest.raises(fb.FaucetError, match="no CUDA device"):
        fb.scan_mem(b">x\nACGT\n", False, False, True, 31, 1, 100, np.zeros(8192, np.uint8), 16, 4)
    with pytest.raises(fb.FaucetError):
        fb.Session(31, 16, 4)


def test_record_layout():
    import faucet_b200 as fb
    assert ctypes.sizeof(fb.JunctionRec) == 32 and fb.REC_DTYPE.itemsize == 32
    assert fb.REC_DTYPE.fields["creation_rank"][1] == 24
