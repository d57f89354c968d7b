"""GPU (-m gpu): GDN_DETERMINISTIC=1 -- fixed-order fp32 reductions (BatchNorm statistics in the convolution epilogues and
the reduce kernels, split-K weight gradients through per-split slabs): two runs of the fused RtoD training step
(trainer.py:696-768), and its CUDA-graph replay, must agree bit for bit.  The library reads the switch once, so the check
runs in a subprocess (tools/check_deterministic.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_training_steps_are_bit_identical_in_deterministic_mode():
    env = dict(os.environ, GDN_DETERMINISTIC="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_deterministic.py"), "10", "4"], capture_output=True,
                       text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0 and "DET-OK" in r.stdout and "deterministic=True" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
