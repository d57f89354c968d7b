"""Library GPU baseline for context: the reference's RtoD training step restated with torch ops
(oracle port: F.conv2d / batch_norm / interpolate -> cuDNN / ATen) on the same B200,
  (i)  fp32 with TF32 disabled            -- the parity oracle's arithmetic on the GPU
  (ii) bf16 autocast + channels_last      -- the strongest off-the-shelf configuration
next to the hand-written path.  TOOLING: imports oracle/ (allowed for tools that measure the baseline, never by the
product).  Not yet run on hardware (written at the end of round 1, when the GPU budget was spent).
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


def make_step(dev, B, bf16):
    from oracle import model as OM, losses as OL, synth
    from tests.util import shapes_of
    sd = {k: v.to(dev) for k, v in synth.synth_state_dict(shapes_of("AutoEncoder_2"), seed=0, bn_random=False).items()}
    sdd = {k: v.to(dev) for k, v in synth.synth_state_dict(shapes_of("AutoEncoder_DtoD"), seed=1, bn_random=False).items()}
    if bf16:
        for d in (sd, sdd):
            for k, v in d.items():
                if v.dim() == 4:
                    d[k] = v.contiguous(memory_format=torch.channels_last)
    pn = [k for k in sd if sd[k].dtype == torch.float32 and not k.endswith(("running_mean", "running_var"))]
    for k in pn:
        sd[k].requires_grad_(True)
    opt = torch.optim.Adam([sd[k] for k in pn], 2e-5, (0.9, 0.999), eps=1e-8, weight_decay=5e-4, fused=True)
    rgb, dep, spa = [t.to(dev) for t in bench.synth_batch(B, 0)]
    if bf16:
        rgb = rgb.contiguous(memory_format=torch.channels_last)

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
            out = OM.autoencoder_2(sd, rgb, istrain=False, train=True, update_running=True)
            with torch.no_grad():      # canonical accounting: encoder + bottleneck only (what the loss consumes)
                ft_tar = OM.autoencoder_dtod(sdd, dep, encoder_only=True)
                ft = OM.autoencoder_dtod(sdd, out.float(), encoder_only=True)
        terms = OL.rtod_loss(out.float(), dep, spa, rgb.float(), [f.float() for f in ft], [f.float() for f in ft_tar])
        opt.zero_grad(set_to_none=True)
        terms["loss"].backward()
        opt.step()
        return terms["loss"]
    return step


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
