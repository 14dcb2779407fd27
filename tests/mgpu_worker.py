"""torchrun worker for tests/test_multi_gpu.py: the exact sharded job on N GPUs vs the oracle on the whole stream.
   torchrun --nproc-per-node N tests/mgpu_worker.py <reads.fq> <out_dir>"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import faucet_b200 as fb
from _oracle import Oracle
from faucet_b200.multi import ShardedJob, TorchComm


def main():
    path, out_dir = sys.argv[1], sys.argv[2]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    fb._lib._check(fb.lib.faucet_gpu_init(local))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    text = open(path, "rb").read()
    res = {"rank": rank, "ok": True, "msg": ""}
    try:
        # (k, j, no_cleaning, share of shard 0 that rank 0 runs in order before the sharded epoch, forced fallback)
        # + read-only walks: 2 = lazy lookups while the keys fit L2 (here: always), 0 = parked lookups
        for (k, j, no_cleaning, prefix_pct, force_abort, lazy) in ((31, 1, 1, 100, 0, 2), (31, 1, 1, 30, 0, 0), (27, 2, 1, 0, 0, 2),
                                                                  (31, 1, 1, 50, 1, 2), (21, 0, 0, 100, 0, 2), (25, 1, 1, 60, 0, 2)):
            _, lt, nh = fb.geometry_from_reads(60000, 30000, 0.04)
            shards = fb.plan_shards(text, True, world)
            a, b = shards[rank]
            cap = max(y - x for x, y in shards) + 1024
            s = fb.Session(k, lt, nh, j=j, max_spacer_dist=100, max_text_bytes=cap)
            fb.set_tuning("shard_force_abort", force_abort)
            fb.set_tuning("dry_lazy", lazy)
            job = ShardedJob(s, TorchComm(torch.device("cuda", local)), prefix_pct=prefix_pct)
            job.setup()
            o = Oracle()
            o1, o2, _ = o.load_two_filters(text, True, k, lt, nh)
            sg, lg = o.geometry_optimal(3000, 0.01), o.geometry_optimal(6000, 0.01)
            ospf, olpf = np.zeros((1 << sg[0]) // 8, np.uint8), np.zeros((1 << lg[0]) // 8, np.uint8)
            gspf, glpf = ospf.copy(), olpf.copy()
            orecs, ost = o.scan(text, True, True, no_cleaning, k, j, 100, o2, lt, nh, ospf, sg, olpf, lg)
            for rep in range(2):  # the job object is reusable (bench steps)
                s.set_text(text[a:b])
                job.load(True)
                g2, _ = s.get_bloom_full()
                assert np.array_equal(g2, o2), f"rank {rank}: bloo2 differs from the single-stream oracle (k={k})"
                gspf[:] = 0
                glpf[:] = 0
                job.scan(True, True, no_cleaning, gspf, sg, glpf, lg)
                want = "serial" if not no_cleaning else "serial (table growth)" if force_abort else "sharded"
                assert job.last_scan["mode"] == want, (job.last_scan, want)
                res.setdefault("scans", []).append(job.last_scan)
                if rank == 0:
                    grecs, gst = s.junctions()
                    assert gst == ost, (gst, ost)
                    for f in ("kmer", "dist", "cov", "linked"):
                        assert np.array_equal(grecs[f], orecs[f]), f
                    assert np.array_equal(gspf, ospf) and np.array_equal(glpf, olpf)
            dist.barrier()
            s.close_peers()
            s.close()
    except Exception as e:  # noqa: BLE001
        import traceback
        res.update(ok=False, msg=traceback.format_exc())
    json.dump(res, open(os.path.join(out_dir, f"rank{rank}.json"), "w"))
    if not res["ok"]:  # the other ranks may be waiting in a collective: let torchrun tear the job down
        sys.stderr.write(res["msg"])
        sys.stderr.flush()
        os._exit(1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
