"""Library GPU baseline for context: the reference's RtoD training step restated with torch ops
(oracle port: F.conv2d / batch_norm / interpolate -> cuDNN / ATen) on the same B200,
  (i)  fp32 with TF32 disabled            -- the parity oracle's arithmetic on the GPU
  (ii) bf16 autocast + channels_last      -- the strongest off-the-shelf configuration
next to the hand-written path.  TOOLING: imports oracle/ (allowed for tools that measure the baseline, never by the
product).  First run: profiles/r02a_bench_torch.jsonl (38.4 img/s fp32, 260.7 img/s bf16 channels_last at batch 20).  bench.py
runs the same two configurations inside its own JSON line (gpu_library_baseline).
    python tools/bench_torch_gpu.py [--batch 20] [--steps 5]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


make_step = bench.torch_gpu_step_fn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=20)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True            # the reference sets it (GDN_main.py:31)
    for name, bf16 in (("fp32 (TF32 off)", False), ("bf16 autocast + channels_last", True)):
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        step = make_step(dev, args.batch, bf16)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps({"impl": "torch-" + ("bf16" if bf16 else "fp32"), "config": name, "batch": args.batch,
                          "ms_per_step": ms, "images_per_s": args.batch / ms * 1e3, "loss": float(loss.detach()),
                          "tflops": bench.GFLOP_RTOD_TRAIN * args.batch / ms / 1e3}), flush=True)
        del step
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
