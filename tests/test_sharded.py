"""N > 1 path: shard assignment + result gather.  CPU: world_size-2 gloo run of the host logic.
GPU (needs >= 2 devices): the same pairs registered on 2 ranks must be bit-identical to 1 rank."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_shard_indices_partition():
    from quatro_b200.sharded import shard_indices
    for n in (0, 1, 7, 8, 2048):
        for w in (1, 2, 3, 8):
            parts = [shard_indices(n, r, w) for r in range(w)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


def _torchrun(nproc, args, port):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tools" / "sharded_check.py"), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)


def test_gather_with_gloo_world2():
    r = _torchrun(2, ["7", "gloo"], 29611)
    assert r.returncode == 0 and "IDENTICAL" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_two_gpus_match_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    r = _torchrun(2, ["6", "nccl"], 29612)
    assert r.returncode == 0 and "IDENTICAL" in r.stdout, r.stdout + r.stderr
