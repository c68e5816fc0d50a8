"""The one collective of the path on real NCCL (needs >= 2 GPUs; the single-GPU driver box skips it, the world-size-2
gloo test covers the host logic on CPU): a 3-round CGLGAN simulation sharded over 2 GPUs must reproduce the
unsharded simulation (tests/nccl_round_check.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_round_over_nccl_matches_single_process(lib):
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300),
           os.path.join(ROOT, "tests", "nccl_round_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "NCCL_CHECK ok" in res.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_mdgan_fegan_flgan_over_nccl_match_single_process(lib):
    """Single-server MD-GAN with its clients dealt over the ranks (replicated generator, all-reduce of the weighted
    dLoss/dXg), FeGAN with the population dealt over the ranks, FL-GAN rounds over a communicator: each against the same
    simulation in one process (tests/nccl_sharded_algos_check.py)."""
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29900 + os.getpid() % 90),
           os.path.join(ROOT, "tests", "nccl_sharded_algos_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "NCCL_ALGOS ok" in res.stdout
