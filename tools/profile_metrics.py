"""The Eigen-metrics and loss kernels alone at the BASELINE shapes, for an `ncu -k regex:...` capture and CUDA-event times:
   python tools/profile_metrics.py
 compute_errors (src/calculate_error.py:10-103): B = 8 at 128x416 (configs[1]) and at 384x1248 (configs[4]);
 RtoD loss (src/trainer.py:705-757): B = 20 at 128x416 (configs[3])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gdn_pytorch_b200 import ops

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)


def timed(fn, reps=20):
    """DEVICE time per call: the calls are captured in a CUDA graph and replayed (timed eagerly, these few-microsecond
    kernels measure the host: ~20 us of Python / ctypes per launch -- round 2's first numbers, 0.16 ms for the seven
    launches of the metrics at B = 8, were exactly that)"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for (b, h, w) in ((8, 128, 416), (8, 384, 1248), (64, 384, 1248)):
    rgb, dep, spa = [t.to(dev) for t in bench.synth_batch(b, 0, h, w)]
    pred = torch.tanh(torch.randn((b, 1, h, w), device=dev))
    ms = timed(lambda: ops.eigen_metrics_device(spa, dep, pred, crop=True))
    algo = 12.0 * b * h * w
    print("eigen_metrics B=%d %dx%d: %.4f ms  (incl. two torch.zeros launches)  algorithmic 12 B/pixel = %.2f MB -> %.1f GB/s"
          % (b, h, w, ms, algo / 1e6, algo / ms / 1e6))
b, h, w = 20, 128, 416
rgb, dep, spa = [t.to(dev) for t in bench.synth_batch(b, 0, h, w)]
out = torch.tanh(torch.randn((b, 1, h, w), device=dev))
kern = ops.LossKernels(dev)
dpre = torch.zeros((b, h, w), device=dev)


def loss():
    kern.absdiff_max(out, dep)
    kern.loss(0, out, dep, spa, rgb, dpre=dpre)


ms = timed(loss)
algo = 28.0 * b * h * w
print("absdiff_max + loss(RtoD) B=%d %dx%d: %.4f ms  algorithmic 28 B/pixel (+8 for the max pass) = %.2f MB -> %.1f GB/s"
      % (b, h, w, ms, algo / 1e6, (algo + 8.0 * b * h * w) / ms / 1e6))
torch.cuda.profiler.start()
for (b2, h2, w2) in ((8, 128, 416), (8, 384, 1248)):
    _, dep2, spa2 = [t.to(dev) for t in bench.synth_batch(b2, 0, h2, w2)]
    pred2 = torch.tanh(torch.randn((b2, 1, h2, w2), device=dev))
    ops.eigen_metrics_device(spa2, dep2, pred2, crop=True)
loss()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
