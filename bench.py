#!/usr/bin/env python
"""bench.py -- k-mers/s of the Faucet hot path (Bloom load + junction scan) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4] [--scaling weak|strong]

A "step" = one pass of the hot path (parse + pass 1 load + pass 2 scan incl. stitch) over one batch of
synthetic reads.  Default workload = BASELINE.json configs[1] (4.6 Mbp genome, 100x, 150 bp paired-end
FASTQ, k=31, --two_hash which is a no-op on the from-reads path).  Prints ONE JSON line (DESIGN.md
"Measurement" explains every field).  oracle/ is used here only as the CPU baseline (cpu_baseline leg,
--impl reference) and, at N > 1, as the checker of a small sharded job run during warm-up.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: genome, coverage, read length, insert, k, estimated_kmers, singletons, flags
    "c1": dict(genome=1_000_000, cov=30, length=100, insert=300, k=31, est=1_000_000, sing=10_000,
               desc="synthetic 1 Mbp random genome, 30x interlaced 100bp paired-end fastq, k=31"),
    "c2": dict(genome=4_600_000, cov=100, length=150, insert=500, k=31, est=4_600_000, sing=1_000_000,
               desc="synthetic E. coli-sized 4.6 Mbp genome, 100x 150bp paired-end, k=31, --two_hash"),
    "c3": dict(genome=64_000_000, cov=50, length=100, insert=300, k=27, est=64_000_000, sing=20_000_000,
               desc="synthetic 64 Mbp (chr20-sized) genome, 50x 100bp paired-end, k=27"),
    # configs[3]: 1 Gbp of reads against a filter sized for 1e9 k-mers: 1 GiB Bloom arrays, the HBM-resident case
    "c4": dict(genome=100_000_000, cov=10, length=100, insert=300, k=31, est=1_000_000_000, sing=200_000_000,
               desc="synthetic 1 Gbp read stream (100 Mbp genome, 10x, 100bp PE), -estimated_kmers 1e9 -singletons 2e8, k=31"),
}
J, MAX_SPACER, FP = 1, 100, 0.04  # faucet defaults: -j 1, -max_spacer_dist 100, -fp 0.04 (src/Faucet.h:14-48)
METRIC = "k-mers/sec (Bloom load + junction scan)"
KERNELS = ("parse", "load_A", "load_B", "scan_flags", "stitch", "stitch_dry", "stitch_verify", "stitch_flow_prep", "shard_copy", "shard_merge")


def gen_dataset(w, seed, pairs=None, tag="", stream=0, genome=None):
    """stream > 0: another read sample of the SAME genome (the shard of rank `stream` in a multi-GPU job)"""
    from _oracle import gen_reads
    d = os.environ.get("FAUCET_BENCH_TMP", "/tmp/faucet_bench")
    os.makedirs(d, exist_ok=True)
    g = genome or w["genome"]
    path = os.path.join(d, f"{g}_{w['cov']}_{w['length']}_{seed}_{stream}{tag}.fq")
    if not os.path.exists(path):
        kw = dict(genome=g, cov=w["cov"], length=w["length"], insert=w["insert"], seed=seed, stream=stream)
        if pairs:
            kw["pairs"] = pairs
        gen_reads(path + ".tmp", **kw)
        os.replace(path + ".tmp", path)
    return path


def config_of(w, k, lt, nh):
    """the SAME dict in both arms (--impl ours / reference): what the metric is quoted on"""
    return {"workload": w["desc"], "k": k, "log2_tai": lt, "n_hash": nh, "j": J, "max_spacer_dist": MAX_SPACER,
            "fastq": True, "paired_ends": True, "no_cleaning": True}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.p = gpu, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.t.join(timeout=2)
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def ncu_traffic(workload, kernel):
    """DRAM bytes per STEP of `kernel` (all its launches of one step) from the committed ncu --set full capture
    (profiles/ncu_traffic.json)"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p)).get(workload, {}).get(kernel, {}).get("bytes_per_step")
    except (OSError, ValueError):
        return None


def survey_bytes(w, n_hash, weight2, r_contained):
    """SURVEY.md section 8(d) sector model, bytes per k-mer: what the REFERENCE's algorithm moves"""
    rho = w["length"] / (w["length"] - w["k"] + 1)
    m = (1 - weight2 ** n_hash) / (1 - weight2)
    a_load = rho + 32 * n_hash * (1 + r_contained)
    a_scan = rho + 32 * (n_hash + 6 * m)
    return a_load, a_scan


# share of shard 0 that GPU 0 stitches in order before the sharded epoch starts (None: ShardedJob's default for the world size)
SHARD_PREFIX_PCT = int(os.environ["FAUCET_SHARD_PREFIX_PCT"]) if os.environ.get("FAUCET_SHARD_PREFIX_PCT") else None
L2_BYTES = 100e6  # what of the 126 MB L2 a randomly probed structure can count on


def kernel_bytes(w, n_hash, lt, r_contained, text_per_kmer, junctions, kmers=1):
    """Algorithmic HBM bytes per k-mer of every kernel of THIS implementation.  Streams are
    counted in full; a randomly probed structure costs one 32-byte DRAM sector per probe -- but only when it does not
    fit in L2 (at configs[0..2] the Bloom arrays, and at configs[0..1] the junction keys, are L2 resident: SURVEY H5).
      parse        the text in, 3/8 of it out (validity + 2-bit planes)
      load_A/B     the planes in; n (1 + r) probes of the fused {bloo1, bloo2} array (2 tai / 8 bytes)
      scan_flags   planes in, one flag byte out per text byte, one probe of the memo (tai / 2 entries of 8 bytes) per k-mer
      stitch       planes + flags in; two probes of the key array per k-mer position (both orientations); per record its
                   slot / predecessor rows (256 B), a done flag and two 32-byte record sectors per landing (~5)
      flow_prep    per record: its code words, ~15 (slot, record) pairs of 8 bytes read and written per radix pass (3), rows
    Returns (bytes per k-mer, {structure: resident in L2?})."""
    kpr = w["length"] - w["k"] + 1  # k-mers per record
    planes = 0.375 * text_per_kmer
    tai = 1 << lt
    cap = 1 << 22
    while cap // 2 < junctions + 1_100_000:
        cap *= 2
    res = {"bloom_fused": 2 * tai / 8 <= L2_BYTES, "scan_memo": max(tai // 2, 1 << 20) * 8 <= L2_BYTES, "junction_keys": cap * 8 <= L2_BYTES}
    sec = lambda resident: 0.0 if resident else 32.0  # noqa: E731
    model = {
        "parse": 1.375 * text_per_kmer,
        "load_A": planes + sec(res["bloom_fused"]) * n_hash * (1 + r_contained),
        "load_B": planes,
        "scan_flags": planes + text_per_kmer + sec(res["scan_memo"]),
        "stitch": planes + text_per_kmer + 2 * sec(res["junction_keys"]) + (256 + 32 + 5 * 64) / kpr,
        "stitch_dry": planes + text_per_kmer + 2 * sec(res["junction_keys"]) + (128 + 5 * 32) / kpr,
        "stitch_verify": 128.0 / kpr,
        "stitch_flow_prep": planes + (15 * 8 * 2 * 3 + 256) / kpr,
        # sharded epoch, per k-mer of ONE shard: the table comes in once (keys + records, written locally) / the owner reads
        # keys + 16 bytes of counts per slot of every replica (bench: up to 8)
        "shard_copy": 2 * 72.0 * cap / max(1, kmers),
        "shard_merge": 24.0 * cap / max(1, kmers),
    }
    return model, res


def cpu_geometry(w):
    """Bloom geometry as the reference derives it, from oracle/_ref (or the C port): no product code in this arm"""
    from _oracle import Oracle, Ref, have_ref
    if have_ref():
        r = Ref()
        p1 = ctypes.c_float(r.lib.ref_brent_p1(w["est"], w["sing"], FP)).value
        return r.geometry_optimal(w["est"], p1)
    o = Oracle()
    p1 = ctypes.c_float(o.lib.fo_brent_p1(w["est"], w["sing"], FP)).value
    return o.geometry_optimal(w["est"], p1)


def cpu_reference_sample(w, lt, nh, repeats=1):
    """Times the reference's own CPU implementation (oracle/_ref: the unmodified sources) -- or, if it was not built,
    the C port in oracle/ -- on a bounded sample that keeps the workload's SHAPE: same coverage, read length, insert
    size, k, Bloom geometry and flags, over a proportionally smaller genome (so that one pass is ~10 M k-mers).
    1 thread: the reference is single-threaded."""
    from _oracle import Oracle, Ref, have_ref
    k = w["k"]
    target_kmers = int(os.environ.get("FAUCET_REF_SAMPLE_KMERS", "12000000"))
    per_read = w["length"] - k + 1
    genome = max(20_000, int(target_kmers / per_read * w["length"] / w["cov"]))
    genome = min(genome, w["genome"])
    sample = gen_dataset(w, seed=1, genome=genome, tag="_cpu")
    text = open(sample, "rb").read()
    kmers = (text.count(b"\n") // 4) * per_read
    times = []
    kind = "reference" if have_ref() else "port"
    for _ in range(repeats):
        t0 = time.perf_counter()
        if kind == "reference":
            r = Ref()
            _, b2 = r.load_two_filters(sample, True, k, lt, nh)
            r.scan(sample, True, True, 1, k, J, MAX_SPACER, b2, lt, nh)
        else:
            o = Oracle()
            _, b2, _ = o.load_two_filters(text, True, k, lt, nh)
            o.scan(text, True, True, 1, k, J, MAX_SPACER, b2, lt, nh)
        times.append(time.perf_counter() - t0)
    desc = (f"{w['cov']}x {w['length']}bp paired-end reads of a {genome} bp genome ({kmers} k-mers per pass): the workload's "
            f"coverage, read shape, k, flags and full-size Bloom geometry on a smaller genome; load_two_filters + scanReads, --no_cleaning")
    return kind, kmers, times, desc


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    k = w["k"]
    lt, nh = cpu_geometry(w)
    kind, kmers, times, sample = cpu_reference_sample(w, lt, nh, repeats=args.steps + args.warmup)
    timed = times[args.warmup:]
    val = kmers * len(timed) / sum(timed)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "k-mers/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(timed) / len(timed),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": config_of(w, k, lt, nh),
        "cpu_baseline": {"value": val, "unit": "k-mers/s", "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def sharded_parity_check(fb, torch, dist, rank, world, local):
    """N > 1: a small stream through the SAME sharded machinery, compared bit for bit with the oracle on the whole stream"""
    import numpy as np
    from _oracle import Oracle, gen_reads
    from faucet_b200.multi import ShardedJob, TorchComm
    d = os.environ.get("FAUCET_BENCH_TMP", "/tmp/faucet_bench")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, f"parity_{world}.fq")
    if rank == 0 and not os.path.exists(path):
        gen_reads(path + ".tmp", genome=60000, cov=40, length=100, insert=300, seed=77, err=0.004, nrate=0.001, repeats=True)
        os.replace(path + ".tmp", path)
    dist.barrier()
    text = open(path, "rb").read()
    k = 31
    _, lt, nh = fb.geometry_from_reads(60000, 30000, FP)
    shards = fb.plan_shards(text, True, world)
    a, b = shards[rank]
    s = fb.Session(k, lt, nh, j=J, max_spacer_dist=MAX_SPACER, max_text_bytes=max(y - x for x, y in shards) + 1024)
    job = ShardedJob(s, TorchComm(torch.device("cuda", local)), sharded_stitch=os.environ.get("FAUCET_SHARD", "1") != "0",
                     prefix_pct=SHARD_PREFIX_PCT)
    job.setup()
    s.set_text(text[a:b])
    job.load(True)
    g2, _ = s.get_bloom_full()
    job.scan(True, True, True)
    ok = True
    o = Oracle()
    _, o2, _ = o.load_two_filters(text, True, k, lt, nh)
    ok &= bool(np.array_equal(g2, o2))
    if rank == 0:
        orecs, ost = o.scan(text, True, True, 1, k, J, MAX_SPACER, o2, lt, nh)
        grecs, gst = s.junctions()
        ok &= gst == ost and all(np.array_equal(grecs[f], orecs[f]) for f in ("kmer", "dist", "cov", "linked"))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    s.close_peers()
    s.close()
    return "ok" if int(flag.item()) == 1 else "FAILED"


def run_ours(args, w):
    import numpy as np
    import torch
    import torch.distributed as dist
    import faucet_b200 as fb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or fb.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: libfaucet_gpu has no CPU path")
    torch.cuda.set_device(local)
    fb._lib._check(fb.lib.faucet_gpu_init(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    for kv in filter(None, os.environ.get("FAUCET_TUNING", "").split(",")):  # experiments: name=value[,name=value]
        name, val = kv.split("=")
        fb.set_tuning(name, int(val))
    k = w["k"]
    _, lt, nh = fb.geometry_from_reads(w["est"], w["sing"], FP)
    parity = sharded_parity_check(fb, torch, dist, rank, world, local) if world > 1 else None
    # One job = N contiguous shards of one read stream, shard g on GPU g; the sharded job is exact (same result as the
    # reference on the concatenated file).  weak: every shard is a full-size read sample of the workload's genome (N x
    # the coverage in total).  strong: the workload's reads are split into N shards (total work fixed).
    pairs = None
    if args.scaling == "strong" and world > 1:
        pairs = int(w["genome"] * w["cov"] / (2.0 * w["length"])) // world
    path = gen_dataset(w, seed=1, stream=rank, pairs=pairs, tag=f"_s{world}" if pairs else "")
    raw = np.fromfile(path, dtype=np.uint8)
    n_text = raw.size
    reads = int(np.count_nonzero(raw == 10)) // 4
    kmers_per_pass = reads * (w["length"] - k + 1)
    host = torch.empty(n_text, dtype=torch.uint8, pin_memory=True)
    host.numpy()[:] = raw
    del raw
    dev = host.cuda()
    cap = torch.tensor([n_text], device="cuda", dtype=torch.int64)
    if world > 1:
        dist.all_reduce(cap, op=dist.ReduceOp.MAX)

    # a device batch holds < 3 GiB of text (32-bit offsets): a bigger single-GPU workload (configs[2] at N = 1) is fed
    # to the stage API in record-aligned parts, all of them resident in HBM
    PART_MAX = 2 << 30
    parts = [(0, n_text)]
    if world == 1 and n_text > PART_MAX:
        parts = fb.plan_shards((host.data_ptr(), n_text), True, (n_text + PART_MAX - 1) // PART_MAX)
    part_cap = max(b - a for a, b in parts)
    sess = fb.Session(k, lt, nh, j=J, max_spacer_dist=MAX_SPACER,
                      max_text_bytes=(int(cap.item()) if len(parts) == 1 else part_cap) + 4096)
    job = None
    if world > 1:
        from faucet_b200.multi import ShardedJob, TorchComm
        # FAUCET_SHARD=0: the serial stitch on GPU 0; FAUCET_SHARD_PREFIX_PCT: share of shard 0 run in order before the epoch
        job = ShardedJob(sess, TorchComm(torch.device("cuda", local)), sharded_stitch=os.environ.get("FAUCET_SHARD", "1") != "0",
                         prefix_pct=SHARD_PREFIX_PCT)
        job.setup()

    def step_parts(base):  # several batches through the stage API (load accumulates; one junction map)
        sess.reset_filters()
        for a, b in parts:
            sess.set_text((base + a, b - a), device=True)
            sess.parse(True)
            sess.load()
        sess.get_bloom(to_host=False)
        sess.stitch_begin(True, True)
        for a, b in parts:
            sess.set_text((base + a, b - a), device=True)
            sess.parse(True)
            sess.scan_flags()
            sess.stitch_batch()
        return 0

    def step_from(src, device):
        if len(parts) > 1:
            return step_parts(src[0])
        sess.set_text(src, device=device)
        if job is None:
            sess.reset_filters()
            sess.parse(True)
            sess.load()
            sess.get_bloom(to_host=False)
            sess.scan_flags()
            # the dependency sort of the stitch is a pure function of the text: it runs on a second stream next to the
            # flagging (memory-bound next to instruction-bound), as pass 1 of the whole-pass entry points prepares it
            sess.flow_prepare(n_recs_resident(), concurrent=True)
            return sess.stitch(True, True)
        job.load(True)
        job.scan(True, True, True)
        return 0

    def n_recs_resident():
        return sess.batch_info()[1]

    def step_resident():
        return step_from((dev.data_ptr(), n_text), True)  # D2D: the batch is already in HBM

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        n_junc = step_resident()
    sess.sync()
    sess.set_profiling(True)
    launches0 = sess.launches
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    sess.timer_start()
    for _ in range(args.steps):
        n_junc = step_resident()
    ms = sess.timer_stop_ms()
    barrier()
    clk = clocks.stop()
    launches = sess.launches - launches0
    kernel_ms = {name: sess.kernel_ms(name) for name in KERNELS}
    sess.set_profiling(False)
    lstats = sess.load_stats()
    stitch_info, n_junc = {}, 0
    if rank == 0:
        recs, _ = sess.junctions()  # gathers the map once (not timed); publishes the stitch counters
        n_junc = len(recs)
        stitch_info = {k_: v for k_, v in fb.timings().items()
                       if k_.startswith(("stitch_r", "stitch_d", "epoch", "exact_", "dry_", "nonquiet", "writer"))}
    b2, _ = sess.get_bloom_full() if world > 1 else sess.get_bloom()
    weight2 = float(np.unpackbits(b2).sum()) / (1 << lt)

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D + D2H inside the timed region
    hptr = (host.data_ptr(), n_text)
    two_uploads = False
    bloo2 = np.empty((1 << lt) // 8, np.uint8)
    e2e_steps = args.steps
    extra = {}
    if world == 1:
        e2e_parts = [0.0, 0.0]
        # One upload for both passes: pass 1 leaves the parsed planes of the stream in HBM ("retain_planes") and
        # pass 2 runs on them (faucet_gpu_scan_retained) -- what src/Faucet.cpp does with one reads file, without
        # reading it twice.  FAUCET_BENCH_TWO_UPLOADS=1 times the text-in-both-calls form instead.
        two_uploads = bool(os.environ.get("FAUCET_BENCH_TWO_UPLOADS"))
        if not two_uploads:
            fb.set_tuning("retain_planes", 1)

        def step_e2e():
            ta = time.perf_counter()
            fb.load_two_filters_mem(hptr, True, k, lt, nh, out=bloo2)
            tb = time.perf_counter()
            if two_uploads:
                recs, st = fb.scan_mem(hptr, True, True, True, k, J, MAX_SPACER, bloo2, lt, nh)
            else:
                recs, st = fb.scan_retained(True, True, k, J, MAX_SPACER, None, lt, nh)
            e2e_parts[0] += tb - ta
            e2e_parts[1] += time.perf_counter() - tb
            return len(recs)
        sess.close()
    else:
        def step_e2e():  # the sharded job fed from pinned host memory; results read back to the host
            step_from(hptr, False)
            sess.get_bloom_full()
            return len(sess.junctions()[0]) if rank == 0 else 0
    step_e2e()
    if world == 1:
        e2e_parts[0] = e2e_parts[1] = 0.0
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        n_junc_e2e = step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    assert n_junc_e2e == n_junc, (n_junc_e2e, n_junc)
    if world == 1 and not os.environ.get("FAUCET_BENCH_SKIP_EXTRAS"):
        # (a) the drop-in entry points proper: reads FILE (page-cache warm) -> outputs on the host
        fsteps = max(1, min(args.steps, 5))

        def step_file():
            fb.load_two_filters(path, True, k, lt, nh, out=bloo2)
            return len(fb.scan_retained(True, True, k, J, MAX_SPACER, None, lt, nh)[0])
        step_file()
        t0 = time.perf_counter()
        for _ in range(fsteps):
            nf = step_file()
        ef = (time.perf_counter() - t0) / fsteps
        assert nf == n_junc
        extra["e2e_file"] = {"value": kmers_per_pass / ef, "unit": "k-mers/s", "steps": fsteps,
                             "api": "faucet_gpu_load_two_filters(path) + faucet_gpu_scan_retained",
                             "h2d_bytes_per_step": n_text, "d2h_bytes_per_step": bloo2.nbytes + 32 * int(n_junc)}
        # (b) once with graph cleaning's inputs: both pair filters (the long one is replayed on the host)
        sl, snh = fb.geometry_optimal(w["est"] // 20, 0.01)
        ll, lnh = fb.geometry_optimal(w["est"] // 10, 0.01)
        spf, lpf = np.zeros((1 << sl) // 8, np.uint8), np.zeros((1 << ll) // 8, np.uint8)
        t0 = time.perf_counter()
        fb.load_two_filters_mem(hptr, True, k, lt, nh, out=bloo2)
        rc_, _ = fb.scan_retained(True, False, k, J, MAX_SPACER, None, lt, nh, spf, (sl, snh), lpf, (ll, lnh))
        ec = time.perf_counter() - t0
        extra["e2e_cleaning"] = {"value": kmers_per_pass / ec, "unit": "k-mers/s", "steps": 1,
                                 "api": "..._mem + faucet_gpu_scan_retained(no_cleaning=0, short + long pair filters)",
                                 "junctions": len(rc_)}
    scan_ranks = None
    if job is not None:
        barrier()
        scan_ranks = [None] * world  # every rank's host-side phase times of its last scan (ms)
        dist.all_gather_object(scan_ranks, (job.last_scan or {}).get("ms", {}))
        sess.close_peers()
        sess.close()

    # max over ranks, whole-job aggregate
    t = torch.tensor([ms, e2e_s * 1e3], device="cuda", dtype=torch.float64)
    tot = torch.tensor([float(kmers_per_pass)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_all, e2e_ms_all = t.tolist()
    kmers_all = tot.item()

    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        r_contained = 1.0 - (lstats.fresh_kmers / lstats.kmers if lstats.kmers else 0.0)
        a_load, a_scan = survey_bytes(w, nh, weight2, r_contained)
        text_per_kmer = n_text / max(1, kmers_per_pass)
        per_step = {n: v[0] / args.steps for n, v in kernel_ms.items()}
        model, l2_resident = kernel_bytes(w, nh, lt, r_contained, text_per_kmer, int(n_junc), kmers_per_pass)
        dom = max(per_step, key=lambda n: per_step[n])        # the kernel with the most time per step
        achieved = kmers_per_pass * model[dom] / (per_step[dom] * 1e-3) / 1e9
        value = kmers_all * args.steps / (ms_all * 1e-3)
        out = {
            "metric": METRIC, "value": value,
            "unit": "k-mers/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_all / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": config_of(w, k, lt, nh),
            "detail": {"reads_per_gpu": reads, "kmers_per_pass_per_gpu": kmers_per_pass, "text_bytes_per_gpu": n_text,
                       "junctions": int(n_junc), "resident_batches": len(parts),
                       "l2": "inputs (%.0f MB text + planes) exceed the 126 MB L2" % (n_text / 1e6),
                       "parallelism": ("%d contiguous shards of one read stream (one per GPU): exact P2P prefix-OR / OR all-reduce "
                                       "of the Bloom filters over NVLink; junction stitch: %s" % (world, (job.last_scan or {}).get("mode", "serial on GPU 0"))) if world > 1 else "single GPU"},
            "e2e": {"value": kmers_all / (e2e_ms_all * 1e-3), "unit": "k-mers/s",
                    "h2d_bytes_per_step": ((2 * n_text + bloo2.nbytes) if two_uploads else n_text) if world == 1 else n_text,
                    "api": ("faucet_gpu_load_two_filters_mem + faucet_gpu_scan_" + ("mem" if two_uploads else "retained")) if world == 1
                           else "faucet_session_* sharded job (faucet_b200/multi.py)",
                    "d2h_bytes_per_step": bloo2.nbytes + 32 * int(n_junc), "steps": e2e_steps,
                    **({"load_call_ms": 1e3 * e2e_parts[0] / e2e_steps, "scan_call_ms": 1e3 * e2e_parts[1] / e2e_steps}
                       if world == 1 else {})},
            **extra,
            "gpu_launches": int(launches),
            "clocks": clk,
            # the kernel with the largest share of the step, against ITS OWN byte model (kernel_bytes): algorithmic bytes per
            # step / that kernel's CUDA-event time per step.  `traffic` = ncu DRAM bytes of that kernel per step.
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(args.workload, dom), "peak_kind": peak_kind,
                         "bytes_per_kmer": model[dom], "ms_per_step": per_step[dom], "share_of_step": per_step[dom] / (ms_all / args.steps),
                         # every kernel the same way, and SURVEY 8(d)'s whole-path figure (the bytes the REFERENCE's algorithm
                         # would move per k-mer, A_load + A_scan, at this k-mer rate per GPU)
                         "per_kernel_frac": {n: (kmers_per_pass * model[n] / (per_step[n] * 1e-3) / 1e9 / peak) if per_step[n] > 0 else None
                                             for n in per_step},
                         "l2_resident": l2_resident,
                         "survey_bytes_per_kmer": a_load + a_scan,
                         "survey_path_frac": (value / world) * (a_load + a_scan) / 1e9 / peak},
            "kernels_ms_per_step": per_step,
            "bloom_weight2": weight2, "contained_fraction": r_contained, "stitch": stitch_info,
        }
        if parity is not None:
            out["parity_check"] = parity
        if job is not None:
            out["stitch_across_gpus"] = dict(job.last_scan or {"mode": "serial"}, ms_by_rank=scan_ranks)
        if args.cpu_baseline:
            kind, ck, times, sample = cpu_reference_sample(w, lt, nh)
            out["cpu_baseline"] = {"value": ck / times[0], "unit": "k-mers/s", "cores": 1, "kind": kind, "sample": sample}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = one full-size read sample per GPU; strong = the workload's reads split over the GPUs")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours" and not os.environ.get("FAUCET_BENCH_PROFILE_RUN"):
        args.warmup = 3
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
