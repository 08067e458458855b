"""Two-rank data-parallel check as a pytest (needs >= 2 GPUs; skipped on the one-GPU box): spawns tests/dist_gpu_check.py
under torchrun — exact mode bit-identical to one GPU, owner-sharded mode within tolerance and replicas bit-identical."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_rank_data_parallel(built):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "dist_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("bit-identical to single GPU") == 18 and r.stdout.count("owner-sharded") == 21 and r.stdout.count("scatter form") == 8 and r.stdout.count("gather form") == 5, r.stdout[-2000:]
    assert r.stdout.count("pull == push") == 3, r.stdout[-2000:]
    assert r.stdout.count("relation-sharded TransR") == 2 and r.stdout.count("save -> restore -> continue") == 1, r.stdout[-2000:]
