"""BASELINE configs[4] on N GPUs: full-resolution (384x1248) RtoD inference, batch 8 per GPU, Eigen metrics with all-reduced
sums ("fused error-metric reduction") -- checked against ONE GPU evaluating the same 8*N images.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/check_infer_ddp.py

Rank r evaluates the batch with seed r.  Checks (rank 0 prints one 'INFER-DDP-OK' line, any failure raises):
  1. the all-reduced metric sums equal the sums rank 0 obtains by evaluating all N batches itself (fp64, 1e-12 relative);
  2. the per-image delta-threshold pixel COUNTS gathered from the ranks are bit-identical to rank 0's own (integers);
  3. the depth maps themselves are bit-identical across GPUs (eval-mode convolutions use no atomics)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from gdn_pytorch_b200 import ops
from gdn_pytorch_b200.trainer import init_distributed_from_env, shutdown_distributed


def main():
    rank, world, dev = init_distributed_from_env()
    assert world > 1, "run under torchrun with >= 2 ranks"
    B = int(os.environ.get("GDN_BATCH", "8"))
    h, w = bench.FULL_H, bench.FULL_W
    rtod, _ = bench.build_models(dev, h, w)
    rtod.eval()

    def evaluate(seed):
        rgb, dep, spa = [t.to(dev) for t in bench.synth_batch(B, seed, h, w)]
        with torch.no_grad():
            out = rtod(rgb, istrain=False)
            out8, counts = ops.eigen_metrics_device(spa, dep, out, crop=True)
        return out.clone(), out8.clone(), counts.clone()

    out, out8, counts = evaluate(rank)
    red = out8.clone()
    dist.all_reduce(red, op=dist.ReduceOp.SUM)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    chk = torch.stack([out.double().sum(), out.double().abs().sum()])
    all_chk = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(all_chk, chk)
    if rank == 0:
        want8 = torch.zeros_like(out8)
        for r in range(world):
            o, o8, c = evaluate(r)
            want8 += o8
            assert torch.equal(c, all_counts[r]), "delta-threshold pixel counts of rank %d differ from a single-GPU run" % r
            mine = torch.stack([o.double().sum(), o.double().abs().sum()])
            assert torch.equal(mine, all_chk[r]), "depth maps of rank %d are not bit-identical to a single-GPU run" % r
        err = ((red - want8).abs() / want8.abs().clamp_min(1e-300)).max().item()
        assert err <= 1e-12, "all-reduced metric sums differ from the single-GPU sums (%g)" % err
        names = ['abs_diff', 'abs_rel', 'sq_rel', 'a1', 'a2', 'a3', 'rmse', 'rmse_log']
        print("INFER-DDP-OK world=%d images=%d at %dx%d  mean metrics %s  (sum error %.1e, counts and depth maps bit-identical)"
              % (world, B * world, h, w, {n: round(float(v) / world, 6) for n, v in zip(names, red)}, err), flush=True)
    shutdown_distributed([])


if __name__ == "__main__":
    main()
