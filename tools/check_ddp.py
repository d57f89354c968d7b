"""Data-parallel RtoD training step on N GPUs (one process per GPU, NCCL): consistency check.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_ddp.py

Checks (rank 0 prints one 'DDP-OK ...' line, any failure raises):
  1. after K steps the parameters are bit-identical on every rank (the all-reduced gradients and Adam agree);
  2. the CUDA-graph step (NCCL collectives captured) and the eager step follow the same trajectory up to the
     run-to-run noise of the fp32 atomics (see the note at LR);
  3. the BerHu threshold is the GLOBAL-batch max (trainer.py:711-720 computes the loss on the gathered batch);
  4. the bucketed all-reduce covers every gradient element exactly once: the reduced buffer equals a plain
     all_reduce(SUM) of the per-shard gradients kept by a debug hook.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from gdn_pytorch_b200.trainer import RtoDTrainStep, init_distributed_from_env


LR = 1e-6   # Train-mode networks at random init amplify any perturbation ~70x per pass (DESIGN.md "Tolerances"): the
            # fp32 shared-memory atomics of the BatchNorm statistics alone make two IDENTICAL runs differ by ~3e-4 in
            # the first loss and ~10 % in the max-norm of the gradient (tools/check_repro.py measures exactly that:
            # eager vs eager == eager vs graph == with / without side streams).  So graph-vs-eager is held to that
            # level -- same loss trajectory within 1 %, parameter drift well below the total movement.


def run(graph, steps, B, rank, dev):
    os.environ["GDN_GRAPH"] = "1" if graph else "0"
    rtod, dtod = bench.build_models(dev)
    st = RtoDTrainStep(rtod, dtod, lr=LR)
    rgb, dep, spa = [t.to(dev) for t in bench.synth_batch(B, rank)]
    losses = []
    for _ in range(steps):
        losses.append(st.step(rgb, dep, spa)["loss"].clone())
    torch.cuda.synchronize()
    return st, [float(v) for v in losses]


def main():
    rank, world, dev = init_distributed_from_env()
    assert world > 1, "run under torchrun with >= 2 ranks"
    B = int(os.environ.get("GDN_BATCH", "2"))
    steps = 4          # 2 eager warm-up steps, capture + first replay, second replay
    st_g, out_g = run(True, steps, B, rank, dev)
    st_e, out_e = run(False, steps, B, rank, dev)
    # 1. identical replicas
    for st in (st_g, st_e):
        flat = st.flat_params
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, flat), "rank %d: parameters diverged from rank 0" % rank
    # 2. graph == eager (up to atomic-order noise)
    d = (st_g.flat_params - st_e.flat_params).abs().max().item()
    md = (st_g.flat_params - st_e.flat_params).abs().mean().item()
    assert d <= 2.01 * LR * steps and md <= 0.3 * LR * steps, "graph vs eager parameters differ: max %g mean %g" % (d, md)
    from gdn_pytorch_b200 import _lib
    det = bool(_lib.lib().gdn_deterministic())
    if det:      # GDN_DETERMINISTIC=1: fixed-order reductions, same update arithmetic in both launch modes -> exact
        # (the reported loss values are fp64 atomic sums of per-CTA partials: compared to 1e-9, they feed no gradient)
        assert d == 0.0 and all(abs(a - b) <= 1e-9 * abs(b) for a, b in zip(out_g, out_e)), \
            "deterministic mode: graph vs eager differ (max %g) %s %s" % (d, out_g, out_e)
    for a, b in zip(out_g, out_e):
        assert abs(a - b) <= 1e-2 * abs(b), ("loss trajectory", out_g, out_e)
    # 3. global BerHu threshold: every rank holds the same max|diff|
    m = st_e.kern.maxabs.clone()
    m0 = m.clone()
    dist.broadcast(m0, src=0)
    assert torch.equal(m, m0), "BerHu threshold is not global"
    # 4. bucketed, overlapped all-reduce == one plain all-reduce of the per-shard gradients
    st_e.debug_keep_local_grad = True
    rgb, dep, spa = [t.to(dev) for t in bench.synth_batch(B, rank)]
    st_e.step(rgb, dep, spa)
    torch.cuda.synchronize()
    want = st_e.local_grad.clone()
    dist.all_reduce(want, op=dist.ReduceOp.SUM)
    got = st_e.eng.flat_grad
    err = (want - got).abs().max().item()
    assert err <= 1e-6 * max(1.0, want.abs().max().item()), "bucketed all-reduce != plain all-reduce (%g)" % err
    assert want.abs().max().item() > 0
    if rank == 0:
        print("DDP-OK world=%d B/rank=%d steps=%d deterministic=%s  |graph-eager| max %.3g mean %.3g (lr %.0e)  loss(graph)=%s loss(eager)=%s" %
              (world, B, steps, det, d, md, LR, ["%.6f" % v for v in out_g], ["%.6f" % v for v in out_e]), flush=True)
    sys.stdout.flush()
    from gdn_pytorch_b200.trainer import shutdown_distributed
    clean = shutdown_distributed([st_g, st_e])
    if rank == 0:
        print("teardown: process group destroyed cleanly = %s" % clean, flush=True)


if __name__ == "__main__":
    main()
