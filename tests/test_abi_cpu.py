"""CPU: the C-ABI library loads without a GPU / driver and exports every symbol include/gdn_b200.h declares;
host-side logic (bucket planning, gloo all-reduce path) works with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gdn_pytorch_b200", "libgdn_b200.so")
HDR = os.path.join(ROOT, "include", "gdn_b200.h")


def declared_symbols():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gdn_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_header_symbols():
    if not os.path.isfile(LIB):
        subprocess.run(["make", "-j8", "-C", ROOT], check=True)
    lib = ctypes.CDLL(LIB)
    syms = declared_symbols()
    assert len(syms) >= 19, syms
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s
    lib.gdn_version.restype = ctypes.c_int
    assert lib.gdn_version() >= 100
    lib.gdn_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.gdn_last_error(), bytes)


def test_invalid_descriptor_is_reported_not_crashed():
    """no GPU needed: argument validation happens before any CUDA call"""
    from gdn_pytorch_b200 import _lib
    L = _lib.lib()
    rc = L.gdn_conv2d(None, None)
    assert rc == -1
    assert b"null" in L.gdn_last_error()
    d = _lib.ConvDesc()
    d.src0 = _lib.Act(1234, 1, 16, 16, 48, 0)   # 48 channels: unsupported
    d.weights = 1234
    d.stride = 1
    d.cout = d.cout_pad = 64
    assert L.gdn_conv2d(ctypes.byref(d), None) == -2
    assert b"multiples of 64" in L.gdn_last_error()


def test_product_refuses_cpu_tensors():
    from tests.util import build_module
    m = build_module("AutoEncoder_2", init_weights=False)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        m(torch.zeros(1, 3, 32, 64))
    from gdn_pytorch_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA fp32"):
        ops.compute_errors(torch.zeros(1, 1, 32, 64), torch.zeros(1, 1, 32, 64), torch.zeros(1, 1, 32, 64))


def test_dataparallel_replicas_are_refused_with_a_pointer_to_the_fused_steps():
    """nn.DataParallel over several GPUs forwards on replicas whose _parameters are empty: refused, not mis-served"""
    from tests.util import build_module
    m = build_module("AutoEncoder_DtoD", init_weights=False)
    rep = m._replicate_for_data_parallel()
    assert rep._is_replica and len(rep._parameters) == 0
    with pytest.raises(RuntimeError, match="one process per GPU"):
        rep(torch.zeros(1, 1, 32, 64))
    blk = m.res64_down1._replicate_for_data_parallel()
    with pytest.raises(RuntimeError, match="DataParallel replica"):
        blk(torch.zeros(1, 64, 32, 64))


def test_bucket_plan_orders_by_readiness():
    from gdn_pytorch_b200.trainer import GradBuckets
    slots = [("a", 0, 1000), ("b", 1000, 3000), ("c", 4000, 500), ("d", 4500, 6000)]
    ready = {"a": 40, "b": 30, "c": 20, "d": 10}
    b = GradBuckets(slots, ready, bucket_bytes=4 * 3500).buckets
    assert [x[:2] for x in sorted(b)] == [(0, 4000), (4000, 10500)]
    assert b[0][2] <= b[1][2]                    # sorted by the op index after which they are final
    assert sum(e - s for s, e, _ in b) == 10500  # every element in exactly one bucket


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from gdn_pytorch_b200.trainer import GradBuckets, allreduce_avg_
rank = int(sys.argv[1]); world = 2
os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = sys.argv[2]
dist.init_process_group("gloo", rank=rank, world_size=world)
g = torch.Generator().manual_seed(rank)
flat = torch.rand(10000, generator=g)
mine = flat.clone()
slots = [("p%%d" %% i, i * 1000, 1000) for i in range(10)]
b = GradBuckets(slots, {"p%%d" %% i: 10 - i for i in range(10)}, bucket_bytes=4 * 2500).buckets
allreduce_avg_(flat, b)
other = torch.rand(10000, generator=torch.Generator().manual_seed(1 - rank))
assert torch.allclose(flat, (mine + other) / 2, atol=1e-7), "bucketed all-reduce mismatch"
mx = torch.tensor([float(rank + 1)])
dist.all_reduce(mx, op=dist.ReduceOp.MAX)
assert mx.item() == 2.0
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_bucketed_allreduce_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER % ROOT)
    port = str(29600 + os.getpid() % 300)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), port], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out.decode()


def test_device_validator_result_host_logic():
    """validate()'s bookkeeping (/root/reference/src/trainer.py:62-87) on the rows the metric kernel fills: plain
    averages, the 'min_errors' pass (rows visited by increasing abs_diff -- same sums, trainer.py:62-85) and the Make3D
    column subset.  Host logic only: the rows are written by hand instead of by the kernel."""
    from gdn_pytorch_b200.validate import DeviceValidator, ERROR_NAMES
    rows = torch.tensor([[3.0, 0.3, 0.03, 0.7, 0.8, 0.9, 5.0, 0.5],
                         [1.0, 0.1, 0.01, 0.9, 0.95, 0.99, 3.0, 0.3],
                         [2.0, 0.2, 0.02, 0.8, 0.9, 0.95, 4.0, 0.4]], dtype=torch.float64)
    for dataset in ("KITTI", "NYU", "Make3D"):
        v = DeviceValidator.__new__(DeviceValidator)
        v.rows, v.n, v.dataset, v.group = torch.zeros((8, 8), dtype=torch.float64), 3, dataset, None
        v.rows[:3] = rows
        avg, mins, names = v.result()
        cols = [0, 1, 2, 6] if dataset == "Make3D" else list(range(8))
        assert names == ERROR_NAMES[dataset] and len(avg) == len(cols) == len(mins)
        for i, c in enumerate(cols):
            assert abs(avg[i] - rows[:, c].mean().item()) < 1e-12
            assert abs(mins[i] - avg[i]) < 1e-12
    v.n = 0
    avg, _, _ = v.result()
    assert avg == [0.0] * 4


_VAL_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from gdn_pytorch_b200.validate import DeviceValidator
rank = int(sys.argv[1])
os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = sys.argv[2]
dist.init_process_group("gloo", rank=rank, world_size=2)
v = DeviceValidator.__new__(DeviceValidator)
v.rows, v.n, v.dataset, v.group = torch.zeros((4, 8), dtype=torch.float64), 2, "KITTI", None
v.rows[0] = float(rank + 1); v.rows[1] = float(10 * (rank + 1))
avg, mins, names = v.result()          # per-batch rows of both ranks are gathered: mean over 4 rows
assert all(abs(a - (1 + 10 + 2 + 20) / 4.0) < 1e-12 for a in avg), avg
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_device_validator_gathers_rows_across_ranks_gloo_world2(tmp_path):
    script = tmp_path / "v.py"
    script.write_text(_VAL_WORKER % ROOT)
    port = str(29900 + os.getpid() % 90)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), port], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out.decode()
