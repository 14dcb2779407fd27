"""Orchestration of the exact sharded two-pass job over N ranks (one process per GPU).

The compute lives behind an ENGINE with the stage methods of faucet_b200.Session (the C ABI's
faucet_session_* multi-GPU API: parse, bloo1_local, prefix_or, load, get_bloom, or_allreduce, scan_flags,
stitch_begin, stitch_batch, import_planes, ...).  This module only sequences those stages and the
cross-process barriers / small exchanges between them (torch.distributed or any object with the same
three calls).  It holds no compute and no fallback.

Algorithm: shard g = g-th contiguous, record-aligned range of
the read stream.
  pass 1   every rank: parse, OR all k-mers of the shard into a local array; barrier;
           bloo1 := exclusive prefix-OR over ranks (peer HBM over NVLink); exact two-filter load of the
           shard; barrier; in-place OR all-reduce of the per-shard bloo2 arrays; barrier.
  pass 2   every rank: scan_flags over its shard (pure).  Then the stitch (the junction map is one sequential
           state: src/ReadScanner.cpp:61-231), in one of two exact forms:
           sharded epoch (default; faucet_b200/csrc/shard.cuh): rank 0 runs the first records of shard 0 through
             the ordered executor; every rank copies that table, classifies ITS OWN records read-only against
             it (quiet records -- the overwhelming majority once the genome is covered a few times -- only
             count coverage, which commutes), the few records that are not quiet run in stream order on every
             replica, the ranks compare notes until no record may have seen a stale table, and rank 0 merges
             the per-rank coverage counts;
           serial (pair filters are fed, or the table would have to grow under the replicas): rank 0 stitches
             shard 0, then pulls the planes + dependency sort of shard 1, 2, ... and stitches them in order.
"""
import struct
import time


class TorchComm:
    """torch.distributed plumbing: barrier + all_gather of small byte strings"""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.device = torch, dist, device
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def barrier(self):
        if self.device is not None:
            self.torch.cuda.synchronize(self.device)
        self.dist.barrier()

    def all_gather_bytes(self, blob):
        t = self.torch.frombuffer(bytearray(blob), dtype=self.torch.uint8)
        if self.device is not None:
            t = t.to(self.device)
        out = self.torch.empty(self.world * t.numel(), dtype=self.torch.uint8, device=t.device)
        self.dist.all_gather_into_tensor(out, t)
        raw = out.cpu().numpy().tobytes()  # one copy back for all ranks' blobs
        return [raw[i * len(blob):(i + 1) * len(blob)] for i in range(self.world)]


class SoloComm:
    rank, world = 0, 1

    def barrier(self):
        pass

    def all_gather_bytes(self, blob):
        return [blob]


PLANES = ("inval", "packed", "flags", "seq_start", "seq_end")


class ShardedJob:
    def __init__(self, engine, comm, sharded_stitch=True, prefix_pct=None):
        """sharded_stitch: take the sharded epoch when the scan allows it; prefix_pct: share of shard 0 that rank 0
        runs through the ordered executor before the epoch starts (the dense start of the stream).  Default: 40 % at
        2 ranks, 55 % at 4, 70 % at 8 -- the table the epoch starts from is staler for every further shard, so a longer
        prefix pays with more of them (measured at configs[1]: DESIGN.md section 6)."""
        self.eng, self.comm = engine, comm
        self.rank, self.world = comm.rank, comm.world
        self.ready = False
        if prefix_pct is None:
            prefix_pct = min(100, 25 + 15 * max(1, (self.world - 1).bit_length()))
        self.sharded_stitch, self.prefix_pct = sharded_stitch, prefix_pct
        self.last_scan = {}

    def setup(self):
        """allocate the exportable buffers and map every peer's (once per session)"""
        e = self.eng
        e.prepare_multi()
        for what in PLANES + ("bloo1_local", "bloom"):
            handles = self.comm.all_gather_bytes(e.export(what))
            e.open_peers(what, handles, self.world, self.rank)
        self.comm.barrier()
        self.ready = True

    def load(self, fastq):
        """pass 1 over the shard already placed with engine.set_text(); leaves the full bloo2 on every rank"""
        assert self.ready
        e = self.eng
        e.prepare_multi()  # zero the shard-wide bit array
        e.parse(fastq)
        e.bloo1_local()
        e.sync()
        self.comm.barrier()
        e.prefix_or()
        e.load()
        e.get_bloom(to_host=False)
        e.sync()
        self.comm.barrier()
        e.or_allreduce()
        e.sync()
        self.comm.barrier()

    def _open(self, names, blank=False):
        """(buffers may have moved since the last scan: exported each time; one exchange for all of them)"""
        e = self.eng
        mine = b"".join(bytes(64) if blank else e.export(what) for what in names)
        blobs = self.comm.all_gather_bytes(mine)
        for i, what in enumerate(names):
            e.open_peers(what, [b[64 * i:64 * i + 64] for b in blobs], self.world, self.rank)

    def scan(self, fastq, paired_ends, no_cleaning, spf=None, spf_geom=(0, 0), lpf=None, lpf_geom=(0, 0)):
        """pass 2; rank 0 ends up holding the junction map (engine.junctions())"""
        e = self.eng
        # (every rank is called with the same arguments: the scan can be sharded unless it feeds pair filters)
        if (self.sharded_stitch and self.world > 1 and hasattr(e, "shard_info")
                and (no_cleaning or (spf is None and lpf is None))):
            return self._scan_sharded(fastq, paired_ends, no_cleaning, spf, spf_geom, lpf, lpf_geom)
        e.scan_flags()
        self.last_scan = {"mode": "serial"}
        return self._scan_serial(fastq, paired_ends, no_cleaning, spf, spf_geom, lpf, lpf_geom, begun=False)

    def _scan_sharded(self, fastq, paired_ends, no_cleaning, spf, spf_geom, lpf, lpf_geom):
        e, rank, world = self.eng, self.rank, self.world
        n_text, n_recs = e.batch_info()
        # every rank holds a junction table; the pair filters (if any) belong to rank 0's serial path
        e.stitch_begin(paired_ends, no_cleaning, *((spf, spf_geom, lpf, lpf_geom) if rank == 0 else (None, (0, 0), None, (0, 0))))
        r0 = n_recs * self.prefix_pct // 100 if rank == 0 else 0
        stat = {"mode": "sharded", "prefix_records": r0, "iterations": 0, "exact": [], "ms": {}}
        self.last_scan = stat
        t_last = [time.perf_counter()]

        def lap(name):  # host wall clock of this rank per phase (every phase ends synchronised with the device)
            t = time.perf_counter()
            stat["ms"][name] = round(stat["ms"].get(name, 0.0) + (t - t_last[0]) * 1e3, 3)
            t_last[0] = t
        # rank 0 flags the records of its prefix first and the rest once the table is on its way to the others
        if rank == 0:
            e.scan_flags(0, r0)
            if r0:
                e.flow_prepare(r0, concurrent=True)  # the dependency sort of the prefix, next to the flagging of its records
        else:
            e.scan_flags()
            if hasattr(e, "shard_rows"):
                e.shard_rows(0)  # pure work, done while rank 0 runs its prefix
        if rank == 0 and r0:
            e.stitch_records(0, r0, False)
            e.sync()
        lap("prefix")
        # the table as the prefix left it (rank 0 makes room in it for what the epoch may create)
        infos = self.comm.all_gather_bytes(e.shard_info(r0, rank == 0))
        lap("info")
        if not all(struct.unpack_from("<QQQIIII", b)[6] for b in infos):  # (e.g. the round-based executor was selected)
            stat["mode"] = "serial (not eligible)"
            if rank == 0:
                e.scan_flags(r0, n_recs)
            return self._scan_serial(fastq, paired_ends, no_cleaning, spf, spf_geom, lpf, lpf_geom, begun="records", r_begin=r0)
        self._open(("tbl_pack", "jslot"))
        lap("open")
        if rank == 0:
            e.scan_flags(r0, n_recs)
        n = e.shard_begin(infos, world, rank, 0)
        lap("classify")
        self._open(("exact_list", "tbl_keys", "cov_delta"))
        lap("open")
        prev, it, grow = -1, 0, 0
        while True:
            both = [struct.unpack("<II", b) for b in self.comm.all_gather_bytes(struct.pack("<II", n, grow))]
            lap("exchange")
            if any(g for _, g in both):  # the table would have to grow under the replicas: serial path from T0
                e.shard_abort()
                stat["mode"] = "serial (table growth)"
                return self._scan_serial(fastq, paired_ends, no_cleaning, spf, spf_geom, lpf, lpf_geom, begun="records", r_begin=r0)
            counts = [c for c, _ in both]
            total = sum(counts)
            stat["exact"].append(total)
            if total == 0 or total == prev:
                break
            grow = e.shard_execute(counts, it)
            lap("execute")
            n = 0 if grow else e.shard_verify()  # (the next list goes to the other buffer: peers may still read this one)
            lap("verify")
            prev, it = total, it + 1
        stat["iterations"] = it
        fin = e.shard_finish()
        lap("finish")
        stats = self.comm.all_gather_bytes(fin)  # (a barrier: every rank's counts are final)
        lap("exchange")
        if rank == 0:
            e.shard_merge(stats)
            e.sync()
        else:
            e.shard_end()
        lap("merge")
        self.comm.barrier()  # peers keep their tables and counts alive until rank 0 has read them
        lap("exchange")
        return [struct.unpack_from("<QQQI", b)[0:4:3] for b in infos]

    def _scan_serial(self, fastq, paired_ends, no_cleaning, spf, spf_geom, lpf, lpf_geom, begun, r_begin=0):
        """begun: False = nothing of the stitch has happened yet; "batch" = stitch_begin was called; "records" = rank 0
        has also run its records [0, r_begin)"""
        e = self.eng
        ahead = hasattr(e, "flow_prepare") and not begun
        if self.rank > 0 and ahead:
            # the dependency sort of the stitch is a pure function of the shard's text: its owner sorts, rank 0 imports
            e.flow_prepare()
        n_text, n_recs = e.batch_info()
        if self.rank == 0:  # shard 0 needs nothing from the others: stitched while they sort
            if not begun:
                e.stitch_begin(paired_ends, no_cleaning, spf, spf_geom, lpf, lpf_geom)
                e.stitch_batch()
            elif begun == "records":
                e.stitch_records(r_begin, n_recs, True)
            else:
                e.stitch_batch()
        e.sync()
        if hasattr(e, "flow_prepare"):
            self._open(("flow_rows", "flow_preds"), blank=not ahead or self.rank == 0)
        infos = [struct.unpack("<QQ", b) for b in self.comm.all_gather_bytes(struct.pack("<QQ", n_text, n_recs))]
        self.comm.barrier()
        if self.rank == 0:
            for r in range(1, self.world):
                e.import_planes(r, infos[r][0], infos[r][1], fastq)
                e.stitch_batch()
            e.sync()
        self.comm.barrier()  # peers keep their planes alive until rank 0 has read them
        return infos
