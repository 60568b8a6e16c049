"""GPU (>= 2 devices): the z-slab multi-GPU engine is bit-identical to the single-GPU engine.  Launches
tests/dist_check.py under torchrun with one rank per GPU (NCCL).  Skipped on a one-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_engine_matches_single_gpu_bitwise():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (found %d)" % ngpu)
    world = 1
    while world * 2 <= min(ngpu, 8):
        world *= 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    assert "DIST_CHECK OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
