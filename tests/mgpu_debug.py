"""diagnostic: which stage of the sharded pass 1 diverges (torchrun --nproc-per-node 2 tests/mgpu_debug.py reads.fq)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import faucet_b200 as fb
from _oracle import Oracle
from faucet_b200.multi import ShardedJob, TorchComm

path = sys.argv[1]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
fb._lib._check(fb.lib.faucet_gpu_init(local))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
text = open(path, "rb").read()
k, j = 31, 1
_, lt, nh = fb.geometry_from_reads(60000, 30000, 0.04)
shards = fb.plan_shards(text, True, world)
a, b = shards[rank]
s = fb.Session(k, lt, nh, j=j, max_spacer_dist=100, max_text_bytes=max(y - x for x, y in shards) + 1024)
comm = TorchComm(torch.device("cuda", local))
job = ShardedJob(s, comm)
job.setup()
o = Oracle()
nb = (1 << lt) // 8
def rd(what):
    import ctypes as C
    out = np.empty(nb, np.uint8)
    ptr = {"b1local": None}
    return out
# expected per-shard quantities
e_all = []
for (x, y) in shards:
    b1, b2, _ = o.load_two_filters(text[x:y], True, k, lt, nh)
    e_all.append(b1 | b2)
prior = np.zeros(nb, np.uint8)
for r in range(rank):
    prior |= e_all[r]
_, e_b2_local, _ = o.load_two_filters(text[a:b], True, k, lt, nh, bloo1=prior.copy())
_, e_b2_full, _ = o.load_two_filters(text, True, k, lt, nh)

s.set_text(text[a:b])
s.prepare_multi(); s.parse(True); s.bloo1_local(); s.sync(); comm.barrier()
s.prefix_or(); s.sync()
g2, g1 = s.get_bloom(want_bloo1=True)   # split of fused right after prefix_or: bloo1 must equal prior, bloo2 empty
print(rank, "prior ok", np.array_equal(g1, prior), "bloo2 empty", not g2.any(), flush=True)
s.load(); g2, g1 = s.get_bloom(want_bloo1=True); s.sync()
print(rank, "local bloo2 ok", np.array_equal(g2, e_b2_local), "diff bits", int(np.unpackbits(g2 ^ e_b2_local).sum()), flush=True)
comm.barrier()
s.or_allreduce(); s.sync(); comm.barrier()
gf, _ = s.get_bloom_full()
print(rank, "full bloo2 ok", np.array_equal(gf, e_b2_full), "diff bits", int(np.unpackbits(gf ^ e_b2_full).sum()),
      "expected-local-OR ok", flush=True)
dist.barrier(); dist.destroy_process_group()
