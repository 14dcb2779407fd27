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
        assert len(handles) == n_ranks and all(h.rstrip(b"\0") in (what.encode(), b"") for h in handles)  # (all zero: nothing to show)

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


class ShardProtocolEngine(OracleEngine):
    """The same CPU model, answering the calls of the sharded epoch (faucet_b200/csrc/shard.cuh) as well.  It holds no
    second implementation of the epoch -- the junction map still comes from the oracle over the whole stream -- but it
    checks what the ORCHESTRATION owes the engine: every rank makes the same calls in the same order with the same
    counts, the loop ends exactly when no list grew, lists alternate between two buffers, and a rank that reports
    "the table must grow" sends every rank down the serial path."""

    def __init__(self, *a, grow_on=None, growth=(5, 3, 0)):
        super().__init__(*a)
        self.grow_on, self.growth, self.log = grow_on, growth, []
        self.n_exact, self.iter, self.epoch_open = 0, 0, False

    def scan_flags(self, r_begin=None, r_end=None):
        self.log.append(("scan_flags", r_begin, r_end))

    def flow_prepare(self, n=None, concurrent=False):
        self.log.append(("flow_prepare", n))

    def shard_rows(self, r_begin):
        self.log.append(("shard_rows", r_begin))

    def stitch_records(self, begin, end, advance):
        self.log.append(("stitch_records", begin, end, bool(advance)))

    def shard_info(self, r_begin, is_owner):
        n_text, n_recs = self.batch_info()
        return struct.pack("<QQQIIII", n_text, 0, 1 << 10, n_recs, r_begin, 1, 1).ljust(512, b"\0")

    def shard_begin(self, infos, n_ranks, my_rank, owner=0):
        assert len(infos) == n_ranks and all(len(b) == 512 for b in infos)
        self.epoch_open, self.iter = True, 0
        self.n_exact = 10 + my_rank
        self.log.append(("shard_begin",))
        return self.n_exact

    def shard_execute(self, counts, it):
        assert self.epoch_open and it == self.iter and len(counts) == self.world and counts[self.rank] == self.n_exact
        self.log.append(("shard_execute", tuple(counts), it))
        self.iter += 1
        return 1 if (self.grow_on == (self.rank, it)) else 0

    def shard_verify(self):
        g = self.growth[min(self.iter - 1, len(self.growth) - 1)]
        self.n_exact += g if self.rank % 2 == 0 else 0  # (lists only ever grow; not every rank's does)
        self.log.append(("shard_verify", self.n_exact))
        return self.n_exact

    def shard_finish(self):
        self.log.append(("shard_finish",))
        return struct.pack("<32Q", *([self.rank] * 32))

    def shard_merge(self, stats_all):
        assert self.rank == 0 and len(stats_all) == self.world
        assert [struct.unpack("<32Q", b)[0] for b in stats_all] == list(range(self.world))
        self.log.append(("shard_merge",))
        self.epoch_open = False
        self.stream = list(self.peer_texts)  # (the epoch covered every shard)

    def shard_end(self):
        self.log.append(("shard_end",))
        self.epoch_open = False

    def shard_abort(self):
        self.log.append(("shard_abort",))
        self.epoch_open = False


def _shard_worker(rank, world, port, path, out_dir, grow_on):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import faucet_b200 as fb
        from faucet_b200.multi import ShardedJob, TorchComm
        text = open(path, "rb").read()
        k, lt, nh, j = 25, 19, 3, 1
        shards = fb.plan_shards(text, True, world)
        a, b = shards[rank]
        eng = ShardProtocolEngine(k, lt, nh, j, 100, grow_on=grow_on)
        eng.set_text(text[a:b])
        eng.peer_texts = [text[x:y] for x, y in shards]
        job = ShardedJob(eng, TorchComm(), prefix_pct=50)
        job.setup()
        job.load(True)
        job.scan(True, True, True)
        import json
        json.dump({"log": eng.log, "scan": job.last_scan}, open(os.path.join(out_dir, f"log_{rank}.json"), "w"))
        if rank == 0:
            recs, st = eng.junctions()
            np.save(os.path.join(out_dir, "recs.npy"), recs)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,grow_on", [(2, None), (3, None), (3, (1, 1))])
def test_sharded_epoch_protocol(tmp_path, world, grow_on):
    import json
    path = gen_reads(str(tmp_path / "r.fq"), genome=8000, cov=12, length=100, insert=300, seed=5, err=0.004, nrate=0.001)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_shard_worker, args=(world, port, path, str(tmp_path), grow_on), nprocs=world, join=True)
    logs = [json.load(open(tmp_path / f"log_{r}.json")) for r in range(world)]
    text = open(path, "rb").read()
    o = Oracle()
    _, b2, _ = o.load_two_filters(text, True, 25, 19, 3)
    recs, _ = o.scan(text, True, True, True, 25, 1, 100, b2, 19, 3)
    assert np.array_equal(np.load(tmp_path / "recs.npy"), recs)  # whichever path was taken, rank 0 covers the whole stream
    calls = [[c[0] for c in lg["log"]] for lg in logs]
    execs = [[c for c in lg["log"] if c[0] == "shard_execute"] for lg in logs]
    assert all(e == execs[0] for e in execs), "every rank must run the same exact set in the same iterations"
    if grow_on is None:
        # growth (5, 3, 0) on the even ranks: the lists grow after iterations 0 and 1 and not after iteration 2 -> three runs
        assert [lg["scan"]["mode"] for lg in logs] == ["sharded"] * world
        assert len(execs[0]) == 3 and logs[0]["scan"]["iterations"] == 3
        assert logs[0]["scan"]["exact"][-1] == logs[0]["scan"]["exact"][-2]
        assert calls[0][-2:] == ["shard_finish", "shard_merge"] and all(c[-2:] == ["shard_finish", "shard_end"] for c in calls[1:])
        # rank 0: flags + sort of the prefix, the ordered prefix, then the flags of the rest; the others: everything, rows ahead
        r0 = [c for c in logs[0]["log"] if c[0] in ("scan_flags", "stitch_records", "flow_prepare")]
        assert r0[0][:2] == ["scan_flags", 0] and r0[1][0] == "flow_prepare" and r0[2][0] == "stitch_records" and r0[3][0] == "scan_flags"
        assert r0[0][2] == r0[1][1] == r0[2][2] == r0[3][1]
        assert all(lg["log"][0] == ["scan_flags", None, None] and lg["log"][1] == ["shard_rows", 0] for lg in logs[1:])
    else:
        assert all(lg["scan"]["mode"] == "serial (table growth)" for lg in logs)
        assert all("shard_abort" in c and "shard_merge" not in c and "shard_finish" not in c for c in calls)
        assert len(execs[0]) == 2  # iteration 1 reported the growth
        # rank 0 then runs the rest of its shard in order, from where its prefix ended
        recs0 = [c for c in logs[0]["log"] if c[0] == "stitch_records"]
        assert recs0[-1][1] == logs[0]["scan"]["prefix_records"] == recs0[0][2] and recs0[-1][3] is True


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
