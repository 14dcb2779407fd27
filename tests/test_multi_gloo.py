"""world_size-2 (and 3) gloo runs of the multi-GPU ORCHESTRATION (faucet_b200/multi.py) on CPU.

The engine here is a CPU model built on the oracle (test infrastructure): it answers the same stage calls
as faucet_b200.Session and moves the Bloom arrays with gloo collectives instead of NVLink peer reads.  What
this pins without a GPU: the shard planner, the sequencing / barriers of ShardedJob (no deadlock, every
rank takes the same path), and -- the important part -- that the sharded pass-1 algorithm
(local OR, exclusive prefix-OR, exact load from that bloo1, OR all-reduce) and the shard-ordered stitch
reproduce the single-stream result bit for bit.  The GPU engine is held to the same check in
tests/test_multi_gpu.py.
"""
import os
import socket
import struct

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _oracle import Oracle, gen_reads


class OracleEngine:
    def __init__(self, k, lt, nh, j, spacer):
        self.o = Oracle()
        self.k, self.lt, self.nh, self.j, self.spacer = k, lt, nh, j, spacer
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.calls = []

    def set_text(self, text):
        self.text = bytes(text)

    def prepare_multi(self):
        self.calls.append("prepare_multi")
        self.b1local = np.zeros((1 << self.lt) // 8, np.uint8)

    def export(self, what):
        return struct.pack("<64s", what.encode())

    def open_peers(self, what, handles, n_ranks, my_rank):
        assert len(handles) == n_ranks and all(h.rstrip(b"\0") == what.encode() for h in handles)

    def parse(self, fastq):
        self.fastq = fastq

    def bloo1_local(self):
        b1, b2, _ = self.o.load_two_filters(self.text, self.fastq, self.k, self.lt, self.nh)
        self.b1local = b1 | b2  # every bit any k-mer of the shard sets

    def sync(self):
        pass

    def prefix_or(self):
        allb = [torch.zeros(self.b1local.size, dtype=torch.uint8) for _ in range(self.world)]
        dist.all_gather(allb, torch.from_numpy(self.b1local.copy()))
        self.prior = np.zeros_like(self.b1local)
        for r in range(self.rank):
            self.prior |= allb[r].numpy()

    def load(self):
        _, self.b2, _ = self.o.load_two_filters(self.text, self.fastq, self.k, self.lt, self.nh, bloo1=self.prior.copy())

    def get_bloom(self, to_host=True):
        return (self.b2 if to_host else None), None

    def or_allreduce(self):
        t = torch.from_numpy(self.b2.copy())
        dist.all_reduce(t, op=dist.ReduceOp.BOR)
        self.b2 = t.numpy()

    def scan_flags(self):
        pass

    def batch_info(self):
        return len(self.text), self.text.count(b"\n")

    def stitch_begin(self, paired, no_cleaning, *a):
        self.paired, self.no_cleaning, self.stream = paired, no_cleaning, [self.text]

    def stitch_batch(self):
        pass

    def import_planes(self, r, n_text, n_recs, fastq):
        self.stream.append(self.peer_texts[r])
        assert len(self.peer_texts[r]) == n_text

    def junctions(self):
        return self.o.scan(b"".join(self.stream), self.fastq, self.paired, self.no_cleaning, self.k, self.j, self.spacer,
                           self.b2, self.lt, self.nh)


def _worker(rank, world, port, path, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import faucet_b200 as fb
        from faucet_b200.multi import ShardedJob, TorchComm
        text = open(path, "rb").read()
        k, lt, nh, j = 25, 19, 3, 1
        shards = fb.plan_shards(text, True, world)
        a, b = shards[rank]
        eng = OracleEngine(k, lt, nh, j, 100)
        eng.set_text(text[a:b])
        eng.peer_texts = [text[x:y] for x, y in shards]  # stands in for the NVLink pull of a peer's planes
        job = ShardedJob(eng, TorchComm())
        job.setup()
        job.load(True)
        job.scan(True, True, True)
        np.save(os.path.join(out_dir, f"b2_{rank}.npy"), eng.b2)
        if rank == 0:
            recs, st = eng.junctions()
            np.save(os.path.join(out_dir, "recs.npy"), recs)
            np.save(os.path.join(out_dir, "shards.npy"), np.array(shards))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_job_equals_single_stream(tmp_path, world):
    path = gen_reads(str(tmp_path / "r.fq"), genome=20000, cov=20, length=100, insert=300, seed=31, err=0.005, nrate=0.002)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, path, str(tmp_path)), nprocs=world, join=True)
    text = open(path, "rb").read()
    o = Oracle()
    _, b2, _ = o.load_two_filters(text, True, 25, 19, 3)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"b2_{r}.npy"), b2), f"rank {r}: sharded pass 1 differs from the single stream"
    recs, _ = o.scan(text, True, True, True, 25, 1, 100, b2, 19, 3)
    assert np.array_equal(np.load(tmp_path / "recs.npy"), recs)
    shards = np.load(tmp_path / "shards.npy")
    assert shards[0][0] == 0 and shards[-1][1] == len(text) and all(shards[i][1] == shards[i + 1][0] for i in range(world - 1))
    for a, _ in shards[1:]:
        assert text[a - 1:a] == b"\n" and text[:a].count(b"\n") % 4 == 0


def test_plan_shards_edges():
    import faucet_b200 as fb
    assert fb.plan_shards(b"", True, 4) == [(0, 0)] * 4
    t = b">a\nACGT\n>b\nAC"   # ragged tail: the last shard takes it
    sh = fb.plan_shards(t, False, 2)
    assert sh[0][0] == 0 and sh[-1][1] == len(t) and sh[0][1] == sh[1][0] and sh[0][1] in (0, 8)
    big = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, b"ACGT" * (1 + i % 7), b"I" * 4 * (1 + i % 7)) for i in range(1000))
    for n in (1, 2, 5, 8):
        sh = fb.plan_shards(big, True, n)
        assert sh[0][0] == 0 and sh[-1][1] == len(big)
        for (a, b), (c, d) in zip(sh, sh[1:]):
            assert b == c and big[:b].count(b"\n") % 4 == 0
        sizes = [b - a for a, b in sh]
        assert max(sizes) - min(sizes) < 200
