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
  pass 2   every rank: scan_flags over its shard (pure) and the dependency sort of its records (pure);
           barrier; rank 0 stitches shard 0, then pulls the planes + sort of shard 1, 2, ... and stitches
           them in stream order (the junction map is one sequential state: src/ReadScanner.cpp:61-231);
           barrier.
"""
import struct


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
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [bytes(o.cpu().numpy().tobytes()) for o in out]


class SoloComm:
    rank, world = 0, 1

    def barrier(self):
        pass

    def all_gather_bytes(self, blob):
        return [blob]


PLANES = ("inval", "packed", "flags", "seq_start", "seq_end")


class ShardedJob:
    def __init__(self, engine, comm):
        self.eng, self.comm = engine, comm
        self.rank, self.world = comm.rank, comm.world
        self.ready = False

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

    def scan(self, fastq, paired_ends, no_cleaning, spf=None, spf_geom=(0, 0), lpf=None, lpf_geom=(0, 0)):
        """pass 2; rank 0 ends up holding the junction map (engine.junctions())"""
        e = self.eng
        e.scan_flags()
        ahead = hasattr(e, "flow_prepare")
        if self.rank > 0 and ahead:
            # the dependency sort of the stitch is a pure function of the shard's text: its owner sorts, rank 0 imports
            e.flow_prepare()
        if self.rank == 0:  # shard 0 needs nothing from the others: stitched while they sort
            e.stitch_begin(paired_ends, no_cleaning, spf, spf_geom, lpf, lpf_geom)
            e.stitch_batch()
        e.sync()
        if ahead:
            for what in ("flow_rows", "flow_preds"):  # (the buffers may have moved since the last scan: exported each time)
                e.open_peers(what, self.comm.all_gather_bytes(e.export(what) if self.rank > 0 else bytes(64)), self.world, self.rank)
        n_text, n_recs = e.batch_info()
        infos = [struct.unpack("<QQ", b) for b in self.comm.all_gather_bytes(struct.pack("<QQ", n_text, n_recs))]
        self.comm.barrier()
        if self.rank == 0:
            for r in range(1, self.world):
                e.import_planes(r, infos[r][0], infos[r][1], fastq)
                e.stitch_batch()
            e.sync()
        self.comm.barrier()  # peers keep their planes alive until rank 0 has read them
        return infos
