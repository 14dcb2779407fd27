"""ctypes binding of libfaucet_gpu.so (include/faucet_gpu.h).  Fails loudly if the library is missing."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("FAUCET_GPU_LIB") or os.path.join(_HERE, "libfaucet_gpu.so")  # override: kernel experiments only


class FaucetError(RuntimeError):
    pass


class JunctionRec(C.Structure):
    _fields_ = [("kmer", C.c_uint64), ("dist", C.c_uint8 * 5), ("cov", C.c_uint8 * 4), ("linked", C.c_uint8 * 5),
                ("pad", C.c_uint8 * 2), ("creation_rank", C.c_uint64)]


REC_DTYPE = np.dtype([("kmer", "<u8"), ("dist", "u1", 5), ("cov", "u1", 4), ("linked", "u1", 5), ("pad", "u1", 2),
                      ("creation_rank", "<u8")])
assert REC_DTYPE.itemsize == C.sizeof(JunctionRec) == 32


class ScanStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_junctions", "nb_jcheck_kmer", "nb_no_juncs", "nb_processed",
                                           "nb_skipped", "reads_no_errors", "reads_processed",
                                           "unambiguous_reads")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class LoadStats(C.Structure):
    _fields_ = [("reads_processed", C.c_uint64), ("unambiguous_reads", C.c_uint64), ("kmers", C.c_uint64),
                ("fresh_kmers", C.c_uint64), ("weight1", C.c_double), ("weight2", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Timings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d_ms", "parse_ms", "load_ms", "scan_ms", "stitch_ms", "d2h_ms", "total_ms")] + \
               [(n, C.c_uint64) for n in ("kernel_launches", "stitch_rounds", "stitch_deferred")] + \
               [("stitch_phase_ns", C.c_uint64 * 8)] + \
               [(n, C.c_uint64) for n in ("epochs_exact", "epochs_classify", "exact_records", "dry_records", "epoch_iterations",
                                          "epoch_fallbacks", "nonquiet_records", "writer_records")]


_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)
_recpp = C.POINTER(C.POINTER(JunctionRec))


def _load():
    if not os.path.exists(_SO):
        raise FaucetError(f"{_SO} is missing: build it with `make -C faucet_b200/csrc` "
                          "(or __graft_entry__.build()); there is no fallback path")
    L = C.CDLL(_SO)
    L.faucet_gpu_last_error.restype = C.c_char_p
    L.faucet_gpu_version.restype = C.c_char_p
    L.faucet_gpu_init.argtypes = [C.c_int]
    L.faucet_geometry_from_reads.argtypes = [C.c_uint64, C.c_uint64, C.c_float, C.POINTER(C.c_double),
                                             C.POINTER(C.c_int), C.POINTER(C.c_int)]
    for f in (L.faucet_geometry_optimal, L.faucet_geometry_2_hash):
        f.argtypes = [C.c_uint64, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.faucet_gpu_load_two_filters.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, _u8p,
                                              C.POINTER(LoadStats)]
    L.faucet_gpu_load_two_filters_mem.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, _u8p,
                                                  _u8p, C.POINTER(LoadStats)]
    scan_tail = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, C.c_int, _u8p, C.c_int, C.c_int,
                 _u8p, C.c_int, C.c_int, _recpp, _u64p, C.POINTER(ScanStats)]
    L.faucet_gpu_scan.argtypes = [C.c_char_p] + scan_tail
    L.faucet_gpu_scan_mem.argtypes = [C.c_void_p, C.c_size_t] + scan_tail
    L.faucet_gpu_scan_retained.argtypes = scan_tail[1:]  # no text and no fastq flag: the retained planes know
    L.faucet_gpu_free.argtypes = [C.c_void_p]
    L.faucet_gpu_query_ext_masks.argtypes = [_u64p, C.c_uint64, C.c_int, C.c_int, _u8p, C.c_int, C.c_int, _u8p]
    L.faucet_gpu_set_batch_bytes.argtypes = [C.c_size_t]
    L.faucet_gpu_set_epoch_limit.argtypes = [C.c_uint64]
    L.faucet_gpu_set_tuning.argtypes = [C.c_char_p, C.c_uint64]
    L.faucet_gpu_get_timings.argtypes = [C.POINTER(Timings)]
    vp = C.c_void_p
    L.faucet_session_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t]
    L.faucet_session_destroy.argtypes = [vp]
    L.faucet_session_set_text.argtypes = [vp, C.c_void_p, C.c_size_t, C.c_int]
    L.faucet_session_reset_filters.argtypes = [vp]
    L.faucet_session_parse.argtypes = [vp, C.c_int]
    L.faucet_session_load.argtypes = [vp]
    L.faucet_session_scan_flags.argtypes = [vp]
    L.faucet_session_scan_flags_records.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.faucet_session_stitch.argtypes = [vp, C.c_int, C.c_int, _u64p]
    L.faucet_session_stitch_begin.argtypes = [vp, C.c_int, C.c_int, _u8p, C.c_int, C.c_int, _u8p, C.c_int, C.c_int]
    L.faucet_session_stitch_batch.argtypes = [vp]
    L.faucet_session_flow_prepare.argtypes = [vp]
    L.faucet_session_flow_prepare_records.argtypes = [vp, C.c_uint32, C.c_int]
    L.faucet_session_get_bloom.argtypes = [vp, _u8p, _u8p]
    L.faucet_session_set_bloom.argtypes = [vp, _u8p]
    L.faucet_session_read_bloom.argtypes = [vp, _u8p]
    L.faucet_session_get_junctions.argtypes = [vp, _recpp, _u64p, C.POINTER(ScanStats)]
    L.faucet_session_sync.argtypes = [vp]
    L.faucet_session_stream.argtypes = [vp]
    L.faucet_session_stream.restype = C.c_void_p
    L.faucet_session_timer_start.argtypes = [vp]
    L.faucet_session_timer_stop_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.faucet_session_kernel_launches.argtypes = [vp]
    L.faucet_session_kernel_launches.restype = C.c_uint64
    L.faucet_session_kernel_ms.argtypes = [vp, C.c_int, C.POINTER(C.c_float), _u64p]
    L.faucet_session_set_profiling.argtypes = [vp, C.c_int]
    L.faucet_session_load_stats.argtypes = [vp, C.POINTER(LoadStats), C.c_uint64]
    for f in (L.faucet_session_prepare_multi, L.faucet_session_close_peers, L.faucet_session_bloo1_local,
              L.faucet_session_prefix_or, L.faucet_session_or_allreduce):
        f.argtypes = [vp]
    L.faucet_session_export.argtypes = [vp, C.c_int, C.c_void_p]
    L.faucet_session_open_peers.argtypes = [vp, C.c_int, C.c_void_p, C.c_int, C.c_int]
    L.faucet_session_import_planes.argtypes = [vp, C.c_int, C.c_size_t, C.c_uint32, C.c_int]
    L.faucet_session_batch_info.argtypes = [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_uint32)]
    L.faucet_host_plan_shards.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, _u64p]
    u32p = C.POINTER(C.c_uint32)
    L.faucet_session_stitch_records.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int]
    L.faucet_session_shard_rows.argtypes = [vp, C.c_uint32]
    L.faucet_session_shard_info.argtypes = [vp, C.c_uint32, C.c_int, C.c_void_p]
    L.faucet_session_shard_begin.argtypes = [vp, C.c_void_p, C.c_int, C.c_int, C.c_int, u32p]
    L.faucet_session_shard_execute.argtypes = [vp, u32p, C.c_int, C.POINTER(C.c_int)]
    L.faucet_session_shard_verify.argtypes = [vp, u32p]
    L.faucet_session_shard_finish.argtypes = [vp, _u64p]
    L.faucet_session_shard_merge.argtypes = [vp, _u64p]
    L.faucet_session_shard_end.argtypes = [vp]
    L.faucet_session_shard_abort.argtypes = [vp]
    return L


lib = _load()


def _check(rc):
    if rc != 0:
        raise FaucetError(f"libfaucet_gpu error {rc}: {lib.faucet_gpu_last_error().decode()}")


def _ptr(a, t=_u8p):
    return None if a is None else a.ctypes.data_as(t)


def device_count():
    return lib.faucet_gpu_device_count()


def set_batch_bytes(n):
    _check(lib.faucet_gpu_set_batch_bytes(n))


def set_epoch_limit(n):
    _check(lib.faucet_gpu_set_epoch_limit(n))


def set_tuning(name, value):
    _check(lib.faucet_gpu_set_tuning(name.encode(), value))


def timings():
    t = Timings()
    _check(lib.faucet_gpu_get_timings(C.byref(t)))
    return {n: (list(getattr(t, n)) if n == "stitch_phase_ns" else getattr(t, n)) for n, _ in t._fields_}


def plan_shards(text, fastq, n_shards):
    """record-aligned, byte-balanced contiguous ranges [(start, end)] of a FASTA/FASTQ text"""
    addr, n, keep = _as_buffer(text)
    offs = (C.c_uint64 * (n_shards + 1))()
    _check(lib.faucet_host_plan_shards(addr, n, int(fastq), n_shards, offs))
    return [(int(offs[i]), int(offs[i + 1])) for i in range(n_shards)]


def geometry_from_reads(estimated_kmers, singletons, fp=0.04):
    """(p1, log2_tai, n_hash) as getBloomFilterFromReads derives them (src/Faucet.cpp:204-219)"""
    p1, a, b = C.c_double(), C.c_int(), C.c_int()
    _check(lib.faucet_geometry_from_reads(estimated_kmers, singletons, fp, C.byref(p1), C.byref(a), C.byref(b)))
    return p1.value, a.value, b.value


def geometry_optimal(items, fp):
    a, b = C.c_int(), C.c_int()
    _check(lib.faucet_geometry_optimal(items, fp, C.byref(a), C.byref(b)))
    return a.value, b.value


def geometry_2_hash(items, fp):
    a, b = C.c_int(), C.c_int()
    _check(lib.faucet_geometry_2_hash(items, fp, C.byref(a), C.byref(b)))
    return a.value, b.value


def _as_buffer(text):
    """bytes / numpy uint8 array / (address, nbytes) -> (address, nbytes, keepalive)"""
    if isinstance(text, tuple):
        return text[0], text[1], None
    if isinstance(text, (bytes, bytearray)):
        arr = np.frombuffer(text, np.uint8)
        return arr.ctypes.data, arr.size, arr
    arr = np.ascontiguousarray(text).view(np.uint8)
    return arr.ctypes.data, arr.size, arr


def load_two_filters_mem(text, fastq, k, log2_tai, n_hash, want_bloo1=False, out=None):
    """pass 1 over FASTA/FASTQ text in host memory -> (bloo2, bloo1 | None, LoadStats)"""
    addr, n, keep = _as_buffer(text)
    nb = (1 << log2_tai) // 8
    b2 = np.empty(nb, np.uint8) if out is None else out
    b1 = np.empty(nb, np.uint8) if want_bloo1 else None
    st = LoadStats()
    _check(lib.faucet_gpu_load_two_filters_mem(addr, n, int(fastq), k, log2_tai, n_hash, _ptr(b2), _ptr(b1),
                                               C.byref(st)))
    return b2, b1, st


def load_two_filters(path, fastq, k, log2_tai, n_hash, want_bloo1=False, out=None):
    nb = (1 << log2_tai) // 8
    b2 = np.empty(nb, np.uint8) if out is None else out
    b1 = np.empty(nb, np.uint8) if want_bloo1 else None
    st = LoadStats()
    _check(lib.faucet_gpu_load_two_filters(path.encode(), int(fastq), k, log2_tai, n_hash, _ptr(b2), _ptr(b1),
                                           C.byref(st)))
    return b2, b1, st


def query_ext_masks(kmers, k, j, bloo2, log2_tai, n_hash):
    """batched getValidJExtension queries: u8 per k-mer, low nibble = Bloom members among the 4 forward extensions,
    high nibble = those that also pass the depth-j check"""
    km = np.ascontiguousarray(np.asarray(kmers, np.uint64))
    out = np.zeros(len(km), np.uint8)
    _check(lib.faucet_gpu_query_ext_masks(km.ctypes.data_as(_u64p), len(km), k, j, _ptr(bloo2), log2_tai, n_hash, _ptr(out)))
    return out


def _take_recs(recs, n):
    arr = np.zeros(n.value, REC_DTYPE)
    if n.value:
        C.memmove(arr.ctypes.data, recs, n.value * REC_DTYPE.itemsize)
    lib.faucet_gpu_free(recs)
    return arr


def scan_mem(text, fastq, paired_ends, no_cleaning, k, j, max_spacer_dist, bloo2, log2_tai, n_hash, spf=None,
             spf_geom=(0, 0), lpf=None, lpf_geom=(0, 0)):
    """pass 2 over text in host memory -> (records sorted by creation rank, stats dict)"""
    addr, n, keep = _as_buffer(text)
    recs, cnt, st = C.POINTER(JunctionRec)(), C.c_uint64(), ScanStats()
    _check(lib.faucet_gpu_scan_mem(addr, n, int(fastq), int(paired_ends), int(no_cleaning), k, j, max_spacer_dist,
                                   _ptr(bloo2), log2_tai, n_hash, _ptr(spf), spf_geom[0], spf_geom[1], _ptr(lpf),
                                   lpf_geom[0], lpf_geom[1], C.byref(recs), C.byref(cnt), C.byref(st)))
    return _take_recs(recs, cnt), st.as_dict()


def scan_retained(paired_ends, no_cleaning, k, j, max_spacer_dist, bloo2, log2_tai, n_hash, spf=None, spf_geom=(0, 0),
                  lpf=None, lpf_geom=(0, 0)):
    """pass 2 over the planes the last load_two_filters[_mem] left in HBM (set_tuning("retain_planes", 1) before it)"""
    recs, cnt, st = C.POINTER(JunctionRec)(), C.c_uint64(), ScanStats()
    _check(lib.faucet_gpu_scan_retained(int(paired_ends), int(no_cleaning), k, j, max_spacer_dist, _ptr(bloo2), log2_tai,
                                        n_hash, _ptr(spf), spf_geom[0], spf_geom[1], _ptr(lpf), lpf_geom[0], lpf_geom[1],
                                        C.byref(recs), C.byref(cnt), C.byref(st)))
    return _take_recs(recs, cnt), st.as_dict()


def scan(path, fastq, paired_ends, no_cleaning, k, j, max_spacer_dist, bloo2, log2_tai, n_hash, spf=None,
         spf_geom=(0, 0), lpf=None, lpf_geom=(0, 0)):
    recs, cnt, st = C.POINTER(JunctionRec)(), C.c_uint64(), ScanStats()
    _check(lib.faucet_gpu_scan(path.encode(), int(fastq), int(paired_ends), int(no_cleaning), k, j, max_spacer_dist,
                               _ptr(bloo2), log2_tai, n_hash, _ptr(spf), spf_geom[0], spf_geom[1], _ptr(lpf),
                               lpf_geom[0], lpf_geom[1], C.byref(recs), C.byref(cnt), C.byref(st)))
    return _take_recs(recs, cnt), st.as_dict()


class Session:
    """device-resident stage API (faucet_session_* in include/faucet_gpu.h)"""
    KERNELS = {"parse": 0, "load_A": 1, "load_B": 2, "scan_flags": 3, "stitch": 4, "stitch_dry": 5, "stitch_verify": 6, "stitch_flow_prep": 7,
               "shard_copy": 8, "shard_merge": 9}

    def __init__(self, k, log2_tai, n_hash, j=1, max_spacer_dist=100, max_text_bytes=1 << 30):
        self.h = C.c_void_p()
        _check(lib.faucet_session_create(C.byref(self.h), k, log2_tai, n_hash, j, max_spacer_dist, max_text_bytes))
        self.k, self.log2_tai, self.n_hash = k, log2_tai, n_hash

    def close(self):
        if self.h:
            lib.faucet_session_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_text(self, text, device=False):
        addr, n, keep = _as_buffer(text)
        _check(lib.faucet_session_set_text(self.h, addr, n, int(device)))

    def reset_filters(self):
        _check(lib.faucet_session_reset_filters(self.h))

    def parse(self, fastq):
        _check(lib.faucet_session_parse(self.h, int(fastq)))

    def load(self):
        _check(lib.faucet_session_load(self.h))

    def scan_flags(self, r_begin=None, r_end=None):
        if r_begin is None:
            _check(lib.faucet_session_scan_flags(self.h))
        else:
            _check(lib.faucet_session_scan_flags_records(self.h, r_begin, r_end))

    def stitch(self, paired_ends, no_cleaning):
        n = C.c_uint64()
        _check(lib.faucet_session_stitch(self.h, int(paired_ends), int(no_cleaning), C.byref(n)))
        return n.value

    def get_bloom(self, want_bloo1=False, to_host=True):
        nb = (1 << self.log2_tai) // 8
        b2 = np.empty(nb, np.uint8) if to_host else None
        b1 = np.empty(nb, np.uint8) if (want_bloo1 and to_host) else None
        _check(lib.faucet_session_get_bloom(self.h, _ptr(b2), _ptr(b1)))
        return b2, b1

    def get_bloom_full(self):
        """the plain bloo2 array as it stands on the device (after an OR all-reduce: the merged filter)"""
        nb = (1 << self.log2_tai) // 8
        b2 = np.empty(nb, np.uint8)
        _check(lib.faucet_session_read_bloom(self.h, _ptr(b2)))
        return b2, None

    def set_bloom(self, bloo2):
        _check(lib.faucet_session_set_bloom(self.h, _ptr(bloo2)))

    def load_stats(self, total_lines=0):
        st = LoadStats()
        _check(lib.faucet_session_load_stats(self.h, C.byref(st), total_lines))
        return st

    def junctions(self):
        recs, cnt, st = C.POINTER(JunctionRec)(), C.c_uint64(), ScanStats()
        _check(lib.faucet_session_get_junctions(self.h, C.byref(recs), C.byref(cnt), C.byref(st)))
        return _take_recs(recs, cnt), st.as_dict()

    def sync(self):
        _check(lib.faucet_session_sync(self.h))

    # ---- multi-GPU stage API (include/faucet_gpu.h, "multi-GPU") ----
    BUFFERS = {"inval": 0, "packed": 1, "flags": 2, "seq_start": 3, "seq_end": 4, "bloo1_local": 5, "bloom": 6, "flow_rows": 7, "flow_preds": 8,
               "tbl_keys": 9, "tbl_recs": 10, "jslot": 11, "tbl_pack": 12, "exact_list": 13, "cov_delta": 14}

    def prepare_multi(self):
        _check(lib.faucet_session_prepare_multi(self.h))

    def export(self, what):
        buf = C.create_string_buffer(64)
        _check(lib.faucet_session_export(self.h, self.BUFFERS[what], buf))
        return buf.raw

    def open_peers(self, what, handles, n_ranks, my_rank):
        blob = b"".join(handles)
        assert len(blob) == 64 * n_ranks
        _check(lib.faucet_session_open_peers(self.h, self.BUFFERS[what], blob, n_ranks, my_rank))

    def close_peers(self):
        _check(lib.faucet_session_close_peers(self.h))

    def bloo1_local(self):
        _check(lib.faucet_session_bloo1_local(self.h))

    def prefix_or(self):
        _check(lib.faucet_session_prefix_or(self.h))

    def or_allreduce(self):
        _check(lib.faucet_session_or_allreduce(self.h))

    def import_planes(self, peer_rank, n_text, n_recs, fastq):
        _check(lib.faucet_session_import_planes(self.h, peer_rank, n_text, n_recs, int(fastq)))

    def batch_info(self):
        n, r = C.c_size_t(), C.c_uint32()
        _check(lib.faucet_session_batch_info(self.h, C.byref(n), C.byref(r)))
        return n.value, r.value

    def stitch_begin(self, paired_ends, no_cleaning, spf=None, spf_geom=(0, 0), lpf=None, lpf_geom=(0, 0)):
        _check(lib.faucet_session_stitch_begin(self.h, int(paired_ends), int(no_cleaning), _ptr(spf), spf_geom[0],
                                               spf_geom[1], _ptr(lpf), lpf_geom[0], lpf_geom[1]))

    def flow_prepare(self, n=None, concurrent=False):
        if n is None:
            _check(lib.faucet_session_flow_prepare(self.h))
        else:
            _check(lib.faucet_session_flow_prepare_records(self.h, n, int(concurrent)))

    def stitch_batch(self):
        _check(lib.faucet_session_stitch_batch(self.h))

    # ---- the stitch across GPUs: the sharded epoch (include/faucet_gpu.h, faucet_b200/csrc/shard.cuh) ----
    SHARD_INFO_BYTES, SHARD_STATS = 512, 32

    def stitch_records(self, begin, end, advance):
        _check(lib.faucet_session_stitch_records(self.h, begin, end, int(advance)))

    def shard_rows(self, r_begin):
        _check(lib.faucet_session_shard_rows(self.h, r_begin))

    def shard_info(self, r_begin, is_owner):
        buf = C.create_string_buffer(self.SHARD_INFO_BYTES)
        _check(lib.faucet_session_shard_info(self.h, r_begin, int(is_owner), buf))
        return buf.raw

    def shard_begin(self, infos, n_ranks, my_rank, owner=0):
        blob, n = b"".join(infos), C.c_uint32()
        assert len(blob) == self.SHARD_INFO_BYTES * n_ranks
        _check(lib.faucet_session_shard_begin(self.h, blob, n_ranks, my_rank, owner, C.byref(n)))
        return n.value

    def shard_execute(self, counts, it):
        arr, grow = (C.c_uint32 * len(counts))(*counts), C.c_int()
        _check(lib.faucet_session_shard_execute(self.h, arr, it, C.byref(grow)))
        return grow.value

    def shard_verify(self):
        n = C.c_uint32()
        _check(lib.faucet_session_shard_verify(self.h, C.byref(n)))
        return n.value

    def shard_finish(self):
        st = (C.c_uint64 * self.SHARD_STATS)()
        _check(lib.faucet_session_shard_finish(self.h, st))
        return bytes(st)

    def shard_merge(self, stats_all):
        blob = b"".join(stats_all)
        arr = (C.c_uint64 * (len(blob) // 8)).from_buffer_copy(blob)
        _check(lib.faucet_session_shard_merge(self.h, arr))

    def shard_end(self):
        _check(lib.faucet_session_shard_end(self.h))

    def shard_abort(self):
        _check(lib.faucet_session_shard_abort(self.h))

    @property
    def stream(self):
        return lib.faucet_session_stream(self.h)

    def timer_start(self):
        _check(lib.faucet_session_timer_start(self.h))

    def timer_stop_ms(self):
        ms = C.c_float()
        _check(lib.faucet_session_timer_stop_ms(self.h, C.byref(ms)))
        return ms.value

    def set_profiling(self, on):
        _check(lib.faucet_session_set_profiling(self.h, int(on)))

    def kernel_ms(self, name):
        ms, n = C.c_float(), C.c_uint64()
        _check(lib.faucet_session_kernel_ms(self.h, self.KERNELS[name], C.byref(ms), C.byref(n)))
        return ms.value, n.value

    @property
    def launches(self):
        return lib.faucet_session_kernel_launches(self.h)
