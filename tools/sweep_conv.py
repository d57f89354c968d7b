"""Every staging variant of gdn_conv2d on the layer shapes that run below 1000 TFLOP/s in the step (development aid).
    python tools/sweep_conv.py [B]
Prints ms / TFLOP/s per (shape, algo word): mode (1 tap-by-tap, 2 halo-resident), J sub-tiles, channel tile, CTA pairs."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gdn_pytorch_b200 import _lib

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 20
L = _lib.lib()
PAIR = 1 << 24
SPLIT = lambda s: s << 25
EW8 = 1 << 28


def bench(d, reps=3):
    """device time per launch: launches captured in a CUDA graph (eager timing of the small layers measures the host)"""
    from gdn_pytorch_b200.engine import graph_time_ms
    if L.gdn_conv2d(C.byref(d), _lib.stream_ptr()) != 0:
        return None
    torch.cuda.synchronize()
    return graph_time_ms(lambda sp: L.gdn_conv2d(C.byref(d), sp), reps)


def make(cin, cout, k, stride, h, w, kind, cin1=0, pad=None):
    """kind: 'train' (fp16 raw + statistics), 'dgrad' (fp32 out + fp32 accumulate), 'eval' (bias, relu, bf16 + fp32 out)"""
    p = (k // 2) if pad is None else pad
    ho, wo = (h + 2 * p - k) // stride + 1, (w + 2 * p - k) // stride + 1
    ph = p if stride == 2 else 0      # stride-2 sources carry a physical (reflection) border
    x = torch.randn((B, h + 2 * ph, w + 2 * ph, cin), device=dev).to(torch.bfloat16)
    x1 = torch.randn((B, h + 2 * ph, w + 2 * ph, cin1), device=dev).to(torch.bfloat16) if cin1 else None
    wt = (torch.randn((k * k, cout, cin + cin1), device=dev) * 0.02).to(torch.bfloat16)
    d = _lib.ConvDesc()
    d.src0 = _lib.Act(x.data_ptr(), B, h, w, cin, ph)
    if x1 is not None:
        d.src1 = _lib.Act(x1.data_ptr(), B, h, w, cin1, ph)
    d.weights = wt.data_ptr()
    d.kh = d.kw = k
    d.stride = stride
    d.off_y = d.off_x = -p
    d.out_h, d.out_w, d.cout, d.cout_pad = ho, wo, cout, cout
    d.dst_h, d.dst_w, d.dst_sy, d.dst_sx = ho, wo, 1, 1
    keep = [x, x1, wt]
    if kind == "train":
        raw = torch.empty((B, ho, wo, cout), dtype=torch.float16, device=dev)
        st = torch.zeros((2, cout), dtype=torch.float64, device=dev)
        d.out_bf16 = _lib.Act(raw.data_ptr(), B, ho, wo, cout, 0)
        d.out16_is_half = 1
        d.stat_sum, d.stat_sqsum = st[0].data_ptr(), st[1].data_ptr()
        keep += [raw, st]
    elif kind == "dgrad":
        o = torch.zeros((B, ho, wo, cout), dtype=torch.float32, device=dev)
        d.out_f32 = o.data_ptr()
        d.resid = o.data_ptr()
        keep += [o]
    else:
        o = torch.zeros((B, ho, wo, cout), dtype=torch.float32, device=dev)
        ob = torch.zeros((B, ho, wo, cout), dtype=torch.bfloat16, device=dev)
        bias = torch.zeros(cout, device=dev)
        d.out_f32, d.bias, d.relu = o.data_ptr(), bias.data_ptr(), 1
        d.out_bf16 = _lib.Act(ob.data_ptr(), B, ho, wo, cout, 0)
        keep += [o, ob, bias]
    ws = torch.empty(4 * B * ho * wo * cout * 4, dtype=torch.uint8, device=dev) if B * ho * wo <= 128 * 148 else None
    if ws is not None:
        d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
        keep.append(ws)
    flops = 2.0 * B * ho * wo * cout * (cin + cin1) * k * k
    return d, flops, keep


def name(a):
    return "%s J%d bn%s%s%s%s" % ({1: "tap ", 2: "halo"}[a & 0xff], (a >> 8) & 0xff, ((a >> 16) & 0xff) * 64 or "max",
                                  " pair" if a & PAIR else "", " splitK%d" % ((a >> 25) & 7) if (a >> 25) & 7 else "",
                                  " ew8" if a & EW8 else "")


ALGOS = []
for bn in (0, 1, 2, 4):
    for pair in (0, PAIR):
        ALGOS += [1 | (bn << 16) | pair] + [2 | (j << 8) | (bn << 16) | pair for j in (1, 2, 4)]
ALGOS += [a | EW8 for a in ALGOS]

SHAPES = [("512->512 k3 8x26 train", (512, 512, 3, 1, 8, 26, "train")),
          ("512->512 k3 8x26 eval", (512, 512, 3, 1, 8, 26, "eval")),
          ("512->512 k3 8x26 dgrad", (512, 512, 3, 1, 8, 26, "dgrad")),
          ("512->512 k3 16x52 train", (512, 512, 3, 1, 16, 52, "train")),
          ("512->512 k3 16x52 dgrad", (512, 512, 3, 1, 16, 52, "dgrad")),
          ("1024->512 k1 16x52 train (concat)", (512, 512, 1, 1, 16, 52, "train", 512)),
          ("128->64 k1 128x416 train (concat)", (64, 64, 1, 1, 128, 416, "train", 64)),
          ("64->64 k1 128x416 dgrad", (64, 64, 1, 1, 128, 416, "dgrad")),
          ("64->128 k7 s2 128x416 train", (64, 128, 7, 2, 128, 416, "train")),
          ("64->128 k4 s2 128x416 eval", (64, 128, 4, 2, 128, 416, "eval", 0, 1)),
          ("256->512 k3 s2 32x104 train", (256, 512, 3, 2, 32, 104, "train")),
          ("512->512 k4 s2 16x52 eval", (512, 512, 4, 2, 16, 52, "eval", 0, 1)),
          ("256->256 k5 32x104 train", (256, 256, 5, 1, 32, 104, "train"))]

def main():
    for title, args in SHAPES:
        d, flops, keep = make(*args)
        rows = []
        for a in ALGOS + ([x | SPLIT(s) for x in ALGOS for s in (2, 4)] if d.workspace else []):
            d.algo = a
            ms = bench(d)
            if ms is not None:
                rows.append((ms, a))
        rows.sort()
        print("== %s  (%.1f GFLOP)" % (title, flops / 1e9))
        for ms, a in rows[:10]:
            print("   %-28s %8.4f ms  %7.0f TFLOP/s" % (name(a), ms, flops / ms / 1e9))
        sys.stdout.flush()
        del keep


if __name__ == "__main__":
    main()
