#!/usr/bin/env python
"""Benchmark of the GDN hot path on B200 (driver contract: one JSON line on stdout from rank 0).

    python bench.py --gpus 1 --steps 10 --warmup 3                      # this implementation
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1      # reference CPU path (oracle port)
    torchrun ... bench.py --gpus N ...                                  # N ranks, batch-sharded (weak scaling)

Workload (BASELINE.json metric "RtoD train imgs/s @128x416", configs[3]): one RtoD training step per "step" --
AutoEncoder_2 forward, two frozen AutoEncoder_DtoD guidance passes, masked BerHu + latent MSE + smoothness loss,
backward, Adam -- batch 20 per GPU, synthetic KITTI-shaped 128x416 inputs, random-init weights (seed 0).
  value : images/s with the step's inputs already resident in HBM (CUDA events, max over ranks)
  e2e   : images/s through the same public step() with pinned-host inputs copied H2D and the loss read back D2H
          inside the timed region
  roofline : the dominant kernel (64->64 k9 implicit-GEMM conv, 34 % of forward FLOPs) timed alone, live
  cpu_baseline : the oracle port of the reference step timed on this box's host cores on a bounded sample
`--workload infer` measures BASELINE configs[1] instead (RtoD inference batch 8 + guidance features + Eigen metrics).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

H, W = 128, 416
GFLOP_RTOD_TRAIN = 1595.2   # per image, canonical: 3 x 414.75 + 2 x 175.48
GFLOP_RTOD_INFER = 414.75 + 175.48
GFLOP_DTOD_TRAIN = 3 * 339.17
FULL_H, FULL_W = 384, 1248   # KITTI 375x1242 padded to multiples of 16 (no reference network accepts 375x1242)

def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = []
        for i, name in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(s[i].lower().startswith("active") for s in self.samples):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


def synth_batch(B, seed, h=H, w=W):
    """synthetic KITTI-shaped inputs: rgb U[-1,1]; dense depth = vertical ramp + smooth blobs in [-1,1];
    sparse = dense with 70 % of the pixels set to exactly -1.0 (the `sparse > -1` validity test of trainer.py:706).
    Same generators and seeds as oracle/synth.py, restated here so that the product arm never imports oracle/."""
    g = torch.Generator().manual_seed(seed + 11)
    rgb = torch.rand((B, 3, h, w), generator=g) * 2 - 1
    g = torch.Generator().manual_seed(seed + 23)
    ys = torch.linspace(1.0, -1.0, h).view(1, 1, h, 1).expand(B, 1, h, w)
    low = torch.rand((B, 1, max(h // 8, 1), max(w // 8, 1)), generator=g) * 2 - 1
    blobs = torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=False)
    dep = (0.6 * ys + 0.5 * blobs).clamp(-1, 1).contiguous()
    g = torch.Generator().manual_seed(seed + 37)
    m = torch.rand(dep.shape, generator=g) < 0.3
    spa = torch.where(m, dep, torch.full_like(dep, -1.0))
    return rgb, dep, spa


def build_models(dev, h=H, w=W):
    import contextlib, io
    from gdn_pytorch_b200 import AE_model_unet as M
    with contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(0)
        rtod = M.AutoEncoder_2(norm="Batch", input_dim=3, height=h, width=w)
        torch.manual_seed(1)
        dtod = M.AutoEncoder_DtoD(norm="Batch", input_dim=1, height=h, width=w)
    return rtod.to(dev), dtod.to(dev).eval()


def time_dominant_conv(dev, B, profile_launches=0):
    """64->64 k9 same conv on (B,128,416): the kernel that dominates the step, timed alone with CUDA events"""
    import ctypes as C
    from gdn_pytorch_b200 import _lib
    x = torch.randn((B, H, W, 64), device=dev).to(torch.bfloat16)
    w = (torch.randn((81, 64, 64), device=dev) * 0.01).to(torch.bfloat16)
    raw = torch.empty((B, H, W, 64), dtype=torch.float16, device=dev)
    st = torch.zeros((2, 64), dtype=torch.float64, device=dev)
    d = _lib.ConvDesc()
    d.src0 = _lib.Act(x.data_ptr(), B, H, W, 64, 0)
    d.weights = w.data_ptr()
    d.kh = d.kw = 9
    d.stride = 1
    d.off_y = d.off_x = -4
    d.out_h, d.out_w, d.cout, d.cout_pad = H, W, 64, 64
    d.dst_h, d.dst_w, d.dst_sy, d.dst_sx = H, W, 1, 1
    d.out_bf16 = _lib.Act(raw.data_ptr(), B, H, W, 64, 0)
    d.out16_is_half = 1
    d.stat_sum, d.stat_sqsum = st[0].data_ptr(), st[1].data_ptr()
    L = _lib.lib()
    from gdn_pytorch_b200.engine import autotune_conv
    algo = autotune_conv(L, d)           # the same on-device variant selection the engine applies to every layer
    for _ in range(3):
        _lib.check(L.gdn_conv2d(C.byref(d), _lib.stream_ptr()), "conv")
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        L.gdn_conv2d(C.byref(d), _lib.stream_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * B * H * W * 64 * 64 * 81
    if profile_launches:            # tools/profile_conv.py under `ncu --profile-from-start off`: the shipped variant only
        torch.cuda.profiler.start()
        for _ in range(profile_launches):
            L.gdn_conv2d(C.byref(d), _lib.stream_ptr())
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    return ms, flops, algo


def cpu_step_fn(B):
    """the reference's RtoD training step restated on CPU (oracle port): returns a callable doing one step"""
    from oracle import model as OM, losses as OL, synth
    from tests.util import shapes_of
    sd = synth.synth_state_dict(shapes_of("AutoEncoder_2"), seed=0, bn_random=False)
    sdd = synth.synth_state_dict(shapes_of("AutoEncoder_DtoD"), seed=1, bn_random=False)
    pn = [k for k in sd if sd[k].dtype == torch.float32 and not k.endswith(("running_mean", "running_var"))]
    for k in pn:
        sd[k].requires_grad_(True)
    opt = torch.optim.Adam([sd[k] for k in pn], 2e-5, (0.9, 0.999), eps=1e-8, weight_decay=5e-4)
    rgb, dep, spa = synth_batch(B, 0)

    def step():
        out = OM.autoencoder_2(sd, rgb, istrain=False, train=True, update_running=True)
        with torch.no_grad():   # the reference runs the WHOLE DtoD network twice (trainer.py:699-703)
            ft_tar = OM.autoencoder_dtod(sdd, dep, istrain=True)[:4]
            ft = OM.autoencoder_dtod(sdd, out, istrain=True)[:4]
        terms = OL.rtod_loss(out, dep, spa, rgb, ft, ft_tar)
        opt.zero_grad()
        terms["loss"].backward()
        opt.step()
        return float(terms["loss"].detach())
    return step


def torch_gpu_step_fn(dev, B, bf16):
    """the same RtoD training step with torch ops on the GPU (oracle port -> cuDNN / ATen): the library baseline of
    BASELINE.md 2b / SURVEY.md 8d.  bf16=False: fp32 with TF32 disabled (the parity oracle's arithmetic);
    bf16=True: autocast + channels_last, the strongest off-the-shelf configuration.  Canonical accounting like the
    product arm: the frozen DtoD passes stop after the encoder (what the loss consumes)."""
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
    rgb, dep, spa = [t.to(dev) for t in synth_batch(B, 0)]
    if bf16:
        rgb = rgb.contiguous(memory_format=torch.channels_last)

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
            out = OM.autoencoder_2(sd, rgb, istrain=False, train=True, update_running=True)
            with torch.no_grad():
                ft_tar = OM.autoencoder_dtod(sdd, dep, encoder_only=True)
                ft = OM.autoencoder_dtod(sdd, out.float(), encoder_only=True)
        terms = OL.rtod_loss(out.float(), dep, spa, rgb.float(), [f.float() for f in ft], [f.float() for f in ft_tar])
        opt.zero_grad(set_to_none=True)
        terms["loss"].backward()
        opt.step()
        return terms["loss"]
    return step


def gpu_library_baseline(dev, B, steps=5):
    """-> {"fp32_notf32": img/s, "bf16_channels_last": img/s, ...}: torch 2.11 / cuDNN on the same GPU, same step, same
    batch, same run (CUDA events, 3 warm-up steps with cudnn.benchmark like the reference, GDN_main.py:31)"""
    res = {"unit": "images/s", "batch": B, "steps": steps,
           "what": "the same RtoD training step with torch ops (F.conv2d / batch_norm / interpolate -> cuDNN, autograd, "
                   "fused torch Adam) via the oracle port of the reference classes; the reference tree is absent on the GPU box"}
    old = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for key, bf16 in (("fp32_notf32", False), ("bf16_channels_last", True)):
            step = torch_gpu_step_fn(dev, B, bf16)
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            res[key] = B / (e0.elapsed_time(e1) / steps / 1e3)
            del step
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return res


def dominant_conv_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel (64->64 k9 conv, B = 20) from
    the committed ncu --set full capture: profiles/dominant_conv_traffic.json, written by tools/ncu_summary.py from the
    .ncu-rep of tools/profile_conv.py.  None when the file is absent."""
    p = os.path.join(ROOT, "profiles", "dominant_conv_traffic.json")
    if not os.path.isfile(p):
        return None, None
    d = json.load(open(p))
    return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"]), d.get("source")


CPU_SAMPLE_BATCH = 4


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = CPU_SAMPLE_BATCH
    step = cpu_step_fn(B)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    v = B / dt
    line = {
        "impl": "reference", "metric": "RtoD train imgs/s @128x416", "value": v, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # same config as the product arm (the workload measured); each CPU step is a BOUNDED SAMPLE of it, stated below
        "config": {"workload": RTOD_TRAIN_WORKLOAD % 20, "batch_per_gpu": 20, "parallelism": "dp%d" % args.gpus,
                   "l2": L2_NOTE},
        "reference_sample": "each timed step processes %d images (not 20): images/s = %d / step time" % (B, B),
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "every step = the whole RtoD training step on a batch of %d of the 20 images (fp32 torch "
                                   "CPU ops, all host threads) via the oracle port of trainer.py:696-768; /root/reference is "
                                   "not present on the GPU box" % B},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


L2_NOTE = "working set per step (GBs of activations) >> 126 MB L2; no explicit flush"
RTOD_TRAIN_WORKLOAD = ("RtoD training step, batch %d per GPU, 128x416 (BASELINE configs[3]): AutoEncoder_2 fwd + 2 frozen "
                       "DtoD encoder passes + loss + bwd + fused Adam")
_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL prints its version banner) write to fd 1; the driver wants exactly ONE JSON line there.
    Everything written to stdout from here on goes to stderr; emit() writes the JSON line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="train", choices=["train", "train_guided", "train_dtod", "infer", "infer_fullres", "demo"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    from gdn_pytorch_b200.trainer import init_distributed_from_env, RtoDTrainStep
    from gdn_pytorch_b200 import ops
    rank, world, dev = init_distributed_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    B = args.batch or (20 if args.workload.startswith("train") else (1 if args.workload == "demo" else 8))
    h, w = (FULL_H, FULL_W) if args.workload == "infer_fullres" else (H, W)
    rgb_h, dep_h, spa_h = [t.pin_memory() for t in synth_batch(B, rank, h, w)]
    rgb, dep, spa = rgb_h.to(dev), dep_h.to(dev), spa_h.to(dev)
    rtod, dtod = build_models(dev, h, w)
    launches = [0]

    if args.workload in ("train", "train_guided"):
        # train_guided: the opt-in paper-faithful variant -- the latent loss back-propagates
        # through the frozen DtoD encoder (its input-gradient convolutions are extra work: + 2 x 175.48 GFLOP/img)
        guided = args.workload == "train_guided"
        stepper = RtoDTrainStep(rtod, dtod, lr=2e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=5e-4, guidance_grad=guided)

        def step_dev():
            return stepper.step(rgb, dep, spa)

        from gdn_pytorch_b200.data import HostBatchPrefetcher
        feed = HostBatchPrefetcher(dev)

        def step_e2e():
            # public API: pinned host batch -> HostBatchPrefetcher (H2D of the NEXT step's inputs on a copy stream, every
            # step, inside the timed region) -> RtoDTrainStep.step -> D2H read of this step's loss
            if feed.head == feed.tail:
                feed.submit(rgb_h, dep_h, spa_h)
            r, d, s = feed.next()
            feed.submit(rgb_h, dep_h, spa_h)
            return float(stepper.step(r, d, s)["loss"])   # D2H read of the loss
        h2d = (rgb_h.numel() + dep_h.numel() + spa_h.numel()) * 4
        d2h = 8
        gflop_img = GFLOP_RTOD_TRAIN + (175.48 if guided else 0.0)
        metric = "RtoD train imgs/s @128x416"
        workload = RTOD_TRAIN_WORKLOAD % B
        if guided:
            metric = "RtoD train (guidance gradient) imgs/s @128x416"
            workload += " + latent-loss gradient through the frozen DtoD encoder (opt-in, SURVEY 8f row 3)"
    elif args.workload == "train_dtod":
        from gdn_pytorch_b200.trainer import DtoDTrainStep
        dtod.train()
        stepper = DtoDTrainStep(dtod, lr=2e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=5e-4)

        def step_dev():
            return stepper.step(dep, spa)

        def step_e2e():
            d = dep_h.to(dev, non_blocking=True)
            s = spa_h.to(dev, non_blocking=True)
            return float(stepper.step(d, s)["loss"])
        h2d = (dep_h.numel() + spa_h.numel()) * 4
        d2h = 8
        gflop_img = GFLOP_DTOD_TRAIN
        metric = "DtoD train imgs/s @128x416"
        workload = ("DtoD training step, batch %d per GPU, 128x416 (BASELINE configs[2]): AutoEncoder_DtoD fwd + BerHu/"
                    "Sobel loss + bwd + fused Adam" % B)
    elif args.workload == "demo":
        # the body of depth_extract.py's loop for one KITTI-sized frame at batch 1 --
        # imresize -> normalise -> AutoEncoder (one CUDA graph) -> imresize back to the original size, 8-bit output
        import contextlib, io
        import numpy as np
        from gdn_pytorch_b200 import AE_model_unet as M
        from gdn_pytorch_b200.demo import DepthExtractor
        with contextlib.redirect_stdout(io.StringIO()):
            torch.manual_seed(0)
            ae = M.AutoEncoder(height=H, width=W).to(dev).eval()
        ex = DepthExtractor(ae)
        frames_h = [torch.from_numpy(np.random.RandomState(rank * 7 + i).randint(0, 256, (375, 1242, 3)).astype(np.uint8))
                    .pin_memory() for i in range(B)]
        frames = [f.to(dev) for f in frames_h]
        host_out = torch.empty((375, 1242), dtype=torch.uint8).pin_memory()

        def step_dev():
            for f in frames:
                ex(f)

        def step_e2e():
            for f in frames_h:
                host_out.copy_(ex(f.to(dev, non_blocking=True)), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        h2d = B * 375 * 1242 * 3
        d2h = B * 375 * 1242
        gflop_img = 683.35
        metric = "demo imgs/s @375x1242 -> 128x416 -> 375x1242, batch 1"
        workload = ("depth_extract.py loop body, %d frame(s) of 375x1242x3 uint8 per step, one at a time: imresize -> "
                    "normalise -> AutoEncoder B=1 (CUDA graph) -> imresize back -> uint8 depth image" % B)
        stepper = None
    else:
        from gdn_pytorch_b200.module_runtime import encoder_features
        rtod.eval()
        full = args.workload == "infer_fullres"

        def infer(r, d, s):
            with torch.no_grad():
                out = rtod(r, istrain=False)
                if not full:
                    encoder_features(dtod, out)
                out8 = ops.eigen_metrics_device(s, d, out, crop=True)[0]
                if world > 1:                       # "fused error-metric reduction": one 64-byte all-reduce
                    dist.all_reduce(out8, op=dist.ReduceOp.SUM)
                    out8 = out8 / world
                return out8

        def step_dev():
            return infer(rgb, dep, spa)

        def step_e2e():
            r = rgb_h.to(dev, non_blocking=True)
            d = dep_h.to(dev, non_blocking=True)
            s = spa_h.to(dev, non_blocking=True)
            return infer(r, d, s).tolist()
        h2d = (rgb_h.numel() + dep_h.numel() + spa_h.numel()) * 4
        d2h = 64
        if full:
            gflop_img = 414.75 * 9.0
            metric = "RtoD infer imgs/s @384x1248"
            workload = ("RtoD inference at KITTI full resolution (375x1242 -> 384x1248), batch %d per GPU + Eigen metrics "
                        "with all-reduced sums (BASELINE configs[4])" % B)
        else:
            gflop_img = GFLOP_RTOD_INFER
            metric = "RtoD infer imgs/s @128x416"
            workload = "RtoD inference batch %d + DtoD guidance features + Eigen metrics, 128x416 (BASELINE configs[1])" % B

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms

    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start()
    ms = timed(step_dev, args.steps)
    clocks = sampler.summary() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    ms_step = ms / args.steps
    value = B * world / (ms_step / 1e3)
    e2e_v = B * world / (ms_e2e / args.steps / 1e3)

    if rank == 0:
        burst, sustained, hbm, src = peaks()
        cms, cfl, calgo = time_dominant_conv(dev, B)
        achieved = cfl / (cms / 1e3) / 1e12
        traffic, traffic_src = dominant_conv_traffic()
        roof = {"bound": "tensor", "kernel": "conv_igemm_kernel<64> (64->64 k9 s1, halo-resident)", "achieved": achieved,
                "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst,
                "traffic": traffic if B == 20 else None, "traffic_source": traffic_src,
                "peak_source": src + " burst bf16 (kernel timed alone)", "ms_per_launch": cms,
                "variant": "algo 0x%x (%s, J=%d%s)" % (calgo, "halo" if (calgo & 0xff) == 2 else "tapbox", (calgo >> 8) & 0xff,
                                                       ", CTA pairs cta_group::2" if calgo & (1 << 24) else ""),
                "step_frac_of_sustained": (gflop_img * B / (ms_step / 1e3) / 1e3) / sustained}
        roof["step_tflops"] = gflop_img * B / (ms_step / 1e3) / 1e3
        cpu = None
        if world == 1 and not args.no_cpu_baseline and args.workload == "train":   # rank 0 at N = 1 only
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            cstep = cpu_step_fn(CPU_SAMPLE_BATCH)
            cstep()
            nrep, t0 = 0, time.perf_counter()
            while nrep < 3 or (time.perf_counter() - t0 < 12.0 and nrep < 8):
                cstep()
                nrep += 1
            dt = (time.perf_counter() - t0) / nrep
            cpu = {"value": CPU_SAMPLE_BATCH / dt, "unit": "images/s", "cores": cores, "kind": "port",
                   "sample": "1 warm-up + %d timed RtoD training steps on a batch of %d of the 20 images (fp32 torch CPU ops "
                             "with all host threads, oracle port of trainer.py:696-768)" % (nrep, CPU_SAMPLE_BATCH)}
        lib = None
        if world == 1 and not args.no_gpu_baseline and args.workload == "train":
            lib = gpu_library_baseline(dev, B)
        eng = getattr(locals().get("stepper", None), "eng", None)
        if args.workload in ("train", "train_guided"):
            per_step = (eng.launches_fwd + eng.launches_bwd + len(eng.pack_ops) + len(eng.pack_ops_bwd) +
                        sum(e.launches_fwd + getattr(e, "launches_bwd", 0) for e in stepper.deng if e is not None) +
                        2 + 4 + 1 + (5 if args.workload == "train_guided" else 0))
        elif args.workload == "demo":
            # per frame: 3 bytescale + <= 4 resize (in) + normalise + network + 3 bytescale + <= 4 resize (out)
            per_step = B * (ex.eng.launches_fwd + 15)
        elif args.workload == "train_dtod":
            per_step = eng.launches_fwd + eng.launches_bwd + len(eng.pack_ops) + len(eng.pack_ops_bwd) + 2 + 1
        else:
            per_step = sum(e.launches_fwd for e in rtod.__dict__.get("_gdn_engines", {}).values()) + \
                sum(e.launches_fwd for e in dtod.__dict__.get("_gdn_engines", {}).values()) + 1
        line = {
            "metric": metric, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload, "batch_per_gpu": B, "parallelism": "dp%d" % world,
                       "l2": L2_NOTE},
            "e2e": {"value": e2e_v, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": per_step * args.steps,
            "clocks": clocks,
            "roofline": roof,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if lib is not None:
            line["gpu_library_baseline"] = lib
        emit(line)
    if world > 1:
        # orderly teardown: the CUDA graph that captured the NCCL all-reduces is released first (close()), then the
        # process group is destroyed; a watchdog falls back to os._exit(0) if that ever hangs
        from gdn_pytorch_b200.trainer import shutdown_distributed
        sys.stdout.flush()
        shutdown_distributed([locals().get("stepper")])


if __name__ == "__main__":
    main()
