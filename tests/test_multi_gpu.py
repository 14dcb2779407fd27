"""N-GPU exact sharded job (C ABI multi-GPU stage API + faucet_b200/multi.py) vs the oracle on the whole stream.
Needs >= 2 CUDA devices: run it with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`."""
import glob
import json
import os
import socket
import subprocess
import sys

import pytest

from _oracle import gen_reads

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_job_on_gpus(tmp_path, world):
    import faucet_b200 as fb
    if fb.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    path = gen_reads(str(tmp_path / "r.fq"), genome=60000, cov=30, length=100, insert=300, seed=41, err=0.005, nrate=0.002,
                     repeats=True)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "mgpu_worker.py"), path, str(tmp_path)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    res = [json.load(open(f)) for f in sorted(glob.glob(str(tmp_path / "rank*.json")))]
    for r in res:
        assert r["ok"], r["msg"]
    assert p.returncode == 0, p.stderr[-3000:]
    assert len(res) == world
    for r in res:
        assert r["ok"], r["msg"]
