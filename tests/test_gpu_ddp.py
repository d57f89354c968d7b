"""Data-parallel step on real GPUs (needs >= 2 devices; the driver's single-GPU test box skips it).  The host-side
bucket logic is covered on CPU by tests/test_abi_cpu.py::test_bucketed_allreduce_gloo_world2."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_rank_training_step_is_consistent():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "check_ddp.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "DDP-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
