"""ctypes access to the CHECKERS: oracle/liboracle.so (our C restatement) and, when it was built,
oracle/_ref/libfaucet_ref.so (the unmodified reference behind oracle/ref_shim.cpp).

Test infrastructure only -- the product package never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


class JunctionRec(C.Structure):
    _fields_ = [("kmer", C.c_uint64), ("dist", C.c_uint8 * 5), ("cov", C.c_uint8 * 4),
                ("linked", C.c_uint8 * 5), ("pad", C.c_uint8 * 2)]


REC_DTYPE = np.dtype([("kmer", "<u8"), ("dist", "u1", 5), ("cov", "u1", 4), ("linked", "u1", 5),
                      ("pad", "u1", 2)])
assert REC_DTYPE.itemsize == C.sizeof(JunctionRec) == 24


class ScanStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_junctions", "nb_jcheck_kmer", "nb_no_juncs", "nb_processed",
                                           "nb_skipped", "reads_no_errors", "reads_processed",
                                           "unambiguous_reads")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class LoadStats(C.Structure):
    _fields_ = [("reads_processed", C.c_uint64), ("unambiguous_reads", C.c_uint64), ("kmers", C.c_uint64),
                ("weight1", C.c_double), ("weight2", C.c_double)]


def build_oracle():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = [os.path.join(ORACLE_DIR, f) for f in ("faucet_oracle.c", "faucet_oracle.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)


def _ptr(a, t=_u8p):
    return None if a is None else a.ctypes.data_as(t)


class Oracle:
    """oracle/faucet_oracle.c"""

    def __init__(self):
        L = self.lib = C.CDLL(build_oracle())
        L.fo_revcomp.restype = C.c_uint64
        L.fo_revcomp.argtypes = [C.c_uint64, C.c_int]
        L.fo_canon.restype = C.c_uint64
        L.fo_canon.argtypes = [C.c_uint64, C.c_int]
        L.fo_seed.restype = C.c_uint64
        L.fo_seed.argtypes = [C.c_int]
        L.fo_old_hash.restype = C.c_uint64
        L.fo_old_hash.argtypes = [C.c_uint64, C.c_int, C.c_int]
        L.fo_first_kmer.argtypes = [C.c_char_p, C.c_int, _u64p]
        L.fo_kmer_string.argtypes = [C.c_uint64, C.c_int, C.c_char_p]
        L.fo_brent_p1.restype = C.c_double
        L.fo_brent_p1.argtypes = [C.c_uint64, C.c_uint64, C.c_float]
        L.fo_geometry_optimal.argtypes = [C.c_uint64, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.fo_geometry_2_hash.argtypes = [C.c_uint64, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.fo_weight.restype = C.c_double
        L.fo_weight.argtypes = [_u8p, C.c_int]
        L.fo_load_two_filters.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, _u8p,
                                          _u8p, C.POINTER(LoadStats)]
        L.fo_scan.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                              _u8p, C.c_int, C.c_int, _u8p, C.c_int, C.c_int, _u8p, C.c_int, C.c_int, _u64p,
                              C.c_size_t, C.POINTER(C.POINTER(JunctionRec)), _u64p, C.POINTER(ScanStats)]
        L.fo_scan_reads.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_int, _u64p,
                                    C.c_size_t, C.POINTER(C.POINTER(JunctionRec)), _u64p,
                                    C.POINTER(ScanStats)]
        L.fo_junction_line.argtypes = [C.POINTER(JunctionRec), C.c_int, C.c_char_p, C.c_size_t]
        L.fo_free.argtypes = [C.c_void_p]

    def first_kmer(self, s, k):
        out = C.c_uint64()
        self.lib.fo_first_kmer(s.encode() if isinstance(s, str) else s, k, C.byref(out))
        return out.value

    def kmer_string(self, kmer, k):
        buf = C.create_string_buffer(40)
        self.lib.fo_kmer_string(kmer, k, buf)
        return buf.value.decode()

    def geometry_optimal(self, est, fp):
        a, b = C.c_int(), C.c_int()
        self.lib.fo_geometry_optimal(est, fp, C.byref(a), C.byref(b))
        return a.value, b.value

    def geometry_2_hash(self, est, fp):
        a, b = C.c_int(), C.c_int()
        self.lib.fo_geometry_2_hash(est, fp, C.byref(a), C.byref(b))
        return a.value, b.value

    def load_two_filters(self, text, fastq, k, log2_tai, n_hash, bloo1=None, bloo2=None):
        nb = (1 << log2_tai) // 8
        b1 = np.zeros(nb, np.uint8) if bloo1 is None else bloo1
        b2 = np.zeros(nb, np.uint8) if bloo2 is None else bloo2
        st = LoadStats()
        self.lib.fo_load_two_filters(text, len(text), int(fastq), k, log2_tai, n_hash, _ptr(b1), _ptr(b2),
                                     C.byref(st))
        return b1, b2, st

    def _take(self, recs, n):
        arr = np.zeros(n.value, REC_DTYPE)
        if n.value:
            C.memmove(arr.ctypes.data, recs, n.value * 24)
        self.lib.fo_free(recs)
        return arr

    def scan(self, text, fastq, paired, no_cleaning, k, j, max_spacer, bloo2, log2_tai, n_hash, spf=None,
             spf_geom=(0, 0), lpf=None, lpf_geom=(0, 0), fake=None):
        recs = C.POINTER(JunctionRec)()
        n = C.c_uint64()
        st = ScanStats()
        fk = None if fake is None else np.ascontiguousarray(np.sort(np.asarray(fake, np.uint64)))
        self.lib.fo_scan(text, len(text), int(fastq), int(paired), int(no_cleaning), k, j, max_spacer,
                         _ptr(bloo2), log2_tai, n_hash, _ptr(spf), spf_geom[0], spf_geom[1], _ptr(lpf),
                         lpf_geom[0], lpf_geom[1], _ptr(fk, _u64p), 0 if fk is None else len(fk),
                         C.byref(recs), C.byref(n), C.byref(st))
        return self._take(recs, n), st.as_dict()

    def scan_reads(self, reads, k, j, max_spacer, fake):
        arr = (C.c_char_p * len(reads))(*[r.encode() for r in reads])
        fk = np.ascontiguousarray(np.sort(np.asarray(fake, np.uint64)))
        recs = C.POINTER(JunctionRec)()
        n = C.c_uint64()
        st = ScanStats()
        self.lib.fo_scan_reads(arr, len(reads), k, j, max_spacer, _ptr(fk, _u64p), len(fk), C.byref(recs),
                               C.byref(n), C.byref(st))
        return self._take(recs, n), st.as_dict()

    def junction_lines(self, recs, k):
        buf = C.create_string_buffer(256)
        out = []
        for i in range(len(recs)):
            r = JunctionRec.from_buffer_copy(recs[i:i + 1].tobytes())
            self.lib.fo_junction_line(C.byref(r), k, buf, 256)
            out.append(buf.value.decode())
        return out


REF_SO = os.path.join(ORACLE_DIR, "_ref", "libfaucet_ref.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "faucet")


def have_ref():
    return os.path.exists(REF_SO)


class Ref:
    """the unmodified reference, through oracle/ref_shim.cpp"""

    def __init__(self):
        L = self.lib = C.CDLL(REF_SO)
        L.ref_set_k.argtypes = [C.c_int]
        L.ref_revcomp.restype = C.c_uint64
        L.ref_revcomp.argtypes = [C.c_uint64]
        L.ref_get_canon.restype = C.c_uint64
        L.ref_get_canon.argtypes = [C.c_uint64]
        L.ref_old_hash.restype = C.c_uint64
        L.ref_old_hash.argtypes = [C.c_int, C.c_uint64, C.c_int]
        L.ref_seed.restype = C.c_uint64
        L.ref_seed.argtypes = [C.c_int]
        L.ref_brent_p1.restype = C.c_double
        L.ref_brent_p1.argtypes = [C.c_uint64, C.c_uint64, C.c_float]
        for f in (L.ref_geometry_optimal, L.ref_geometry_2_hash):
            f.argtypes = [C.c_uint64, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_int), _u64p]
        L.ref_load_two_filters.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, _u8p, _u8p]
        L.ref_scan.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, C.c_int,
                               _u8p, C.c_int, C.c_int, _u8p, C.c_int, C.c_int, C.POINTER(JunctionRec),
                               C.c_uint64, C.POINTER(ScanStats), C.c_char_p]
        L.ref_scan_fake.argtypes = [C.POINTER(C.c_char_p), C.c_int, _u64p, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(JunctionRec), C.c_uint64, C.POINTER(ScanStats)]

    def set_k(self, k):
        self.lib.ref_set_k(k)

    def iteration_order(self, keys):
        """iteration order of the reference's unordered_map after inserting `keys` in the given order"""
        k = np.ascontiguousarray(np.asarray(keys, np.uint64))
        out = np.zeros(len(k), np.uint64)
        self.lib.ref_iteration_order.argtypes = [_u64p, C.c_uint64, _u64p]
        self.lib.ref_iteration_order(_ptr(k, _u64p), len(k), _ptr(out, _u64p))
        return out

    def valid_j_extension(self, kmers, k, j, bloo2, log2_tai, n_hash):
        """JunctionMap::getValidJExtension for each oriented k-mer"""
        self.set_k(k)
        km = np.ascontiguousarray(np.asarray(kmers, np.uint64))
        out = np.zeros(len(km), np.int32)
        self.lib.ref_valid_j_extension.argtypes = [_u64p, C.c_uint64, C.c_int, _u8p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        self.lib.ref_valid_j_extension(_ptr(km, _u64p), len(km), j, _ptr(bloo2), log2_tai, n_hash,
                                       out.ctypes.data_as(C.POINTER(C.c_int)))
        return out

    def geometry_optimal(self, est, fp):
        a, b, c = C.c_int(), C.c_int(), C.c_uint64()
        self.lib.ref_geometry_optimal(est, fp, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value

    def geometry_2_hash(self, est, fp):
        a, b, c = C.c_int(), C.c_int(), C.c_uint64()
        self.lib.ref_geometry_2_hash(est, fp, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value

    def load_two_filters(self, path, fastq, k, log2_tai, n_hash):
        self.set_k(k)
        nb = (1 << log2_tai) // 8
        b1, b2 = np.zeros(nb, np.uint8), np.zeros(nb, np.uint8)
        self.lib.ref_load_two_filters(path.encode(), int(fastq), log2_tai, n_hash, _ptr(b1), _ptr(b2))
        return b1, b2

    def scan(self, path, fastq, paired, no_cleaning, k, j, max_spacer, bloo2, log2_tai, n_hash, spf=None,
             spf_geom=(0, 0), lpf=None, lpf_geom=(0, 0), cap=1 << 22, junctions_path=None):
        self.set_k(k)
        recs = np.zeros(cap, REC_DTYPE)
        st = ScanStats()
        self.lib.ref_scan(path.encode(), int(fastq), int(paired), int(no_cleaning), j, max_spacer, _ptr(bloo2),
                          log2_tai, n_hash, _ptr(spf), spf_geom[0], spf_geom[1], _ptr(lpf), lpf_geom[0],
                          lpf_geom[1], recs.ctypes.data_as(C.POINTER(JunctionRec)), cap, C.byref(st),
                          None if junctions_path is None else junctions_path.encode())
        assert st.n_junctions <= cap
        return recs[:st.n_junctions].copy(), st.as_dict()

    def scan_fake(self, reads, k, j, max_spacer, fake, cap=4096):
        self.set_k(k)
        arr = (C.c_char_p * len(reads))(*[r.encode() for r in reads])
        fk = np.ascontiguousarray(np.asarray(fake, np.uint64))
        recs = np.zeros(cap, REC_DTYPE)
        st = ScanStats()
        self.lib.ref_scan_fake(arr, len(reads), _ptr(fk, _u64p), len(fk), j, max_spacer,
                               recs.ctypes.data_as(C.POINTER(JunctionRec)), cap, C.byref(st))
        return recs[:st.n_junctions].copy(), st.as_dict()


def sort_recs(recs):
    """canonical order for set comparison (the reference's own order is unordered_map iteration order)"""
    r = recs.copy()
    r["pad"] = 0
    return r[np.argsort(r["kmer"], kind="stable")]


def gen_reads(path, **kw):
    """run tools/gen_reads (built on demand)"""
    exe = os.path.join(ROOT, "tools", "gen_reads")
    src = exe + ".c"
    if not os.path.exists(exe) or os.path.getmtime(src) > os.path.getmtime(exe):
        subprocess.check_call(["gcc", "-O2", "-o", exe, src])
    flags = {"genome": "-g", "cov": "-c", "length": "-l", "insert": "-i", "seed": "-s", "err": "-e",
             "nrate": "-n", "pairs": "-p", "stream": "-S"}
    cmd = [exe, "-o", path]
    for k, v in kw.items():
        if k in flags:
            cmd += [flags[k], str(v)]
        elif k == "repeats" and v:
            cmd.append("-r")
        elif k == "fasta" and v:
            cmd.append("-a")
        elif k == "lower" and v:
            cmd.append("-w")
    subprocess.check_call(cmd)
    return path
