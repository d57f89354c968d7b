"""GPU check of gdn_conv2d_wgrad against torch autograd (fp32).  Usage: python tools/check_wgrad.py [--time]"""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from gdn_pytorch_b200 import _lib

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"


def act_of(t, pad):
    n, hp, wp, c = t.shape
    return _lib.Act(t.data_ptr(), n, hp - 2 * pad, wp - 2 * pad, c, pad)


def run(name, N, H, W, cin, cout, k, stride=1, pad_mode="zero", cin2=0, timing=False):
    g = torch.Generator(device="cpu").manual_seed(len(name))
    p = k // 2 if k != 4 else 1
    x = (torch.rand((N, cin + cin2, H, W), generator=g) * 2 - 1).to(dev).to(torch.bfloat16).float()
    x = x.double()
    w = torch.zeros((cout, cin + cin2, k, k), device=dev, requires_grad=True, dtype=torch.float64)
    if pad_mode == "reflect":
        xin = F.pad(x, (p,) * 4, mode="reflect")
        y = F.conv2d(xin, w, None, stride, 0)
        bufpad = p
    else:
        xin = x
        y = F.conv2d(xin, w, None, stride, p)
        bufpad = 0
    OH, OW = y.shape[2], y.shape[3]
    dy = (torch.rand((N, cout, OH, OW), generator=g) * 2 - 1).to(dev).to(torch.bfloat16).float()
    (ref,) = torch.autograd.grad(y, w, dy.double())
    ref = ref.float()
    ref = ref.permute(2, 3, 1, 0).reshape(k * k, cin + cin2, cout)
    xbuf = xin.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    dyb = dy.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    d = _lib.WgradDesc()
    if cin2:
        x0 = xbuf[..., :cin].contiguous()
        x1 = xbuf[..., cin:].contiguous()
        d.x0, d.x1 = act_of(x0, bufpad), act_of(x1, bufpad)
    else:
        d.x0 = act_of(xbuf, bufpad)
    d.dy = act_of(dyb, 0)
    dw = torch.zeros((k * k, cin + cin2, cout), device=dev)
    d.dw = dw.data_ptr()
    d.kh = d.kw = k
    d.stride = stride
    d.off_y = d.off_x = -p
    d.out_h, d.out_w = OH, OW
    d.cout_pad = cout
    L = _lib.lib()
    _lib.check(L.gdn_conv2d_wgrad(C.byref(d), _lib.stream_ptr()), name)
    torch.cuda.synchronize()
    err = ((dw - ref).abs().max() / ref.abs().max()).item()
    ok = err < 2e-3 and not torch.isnan(dw).any().item()
    msg = "%-36s err=%.2e" % (name, err)
    if timing:
        for _ in range(2):
            L.gdn_conv2d_wgrad(C.byref(d), _lib.stream_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            L.gdn_conv2d_wgrad(C.byref(d), _lib.stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = 2.0 * N * OH * OW * cout * (cin + cin2) * k * k
        msg += "  %.3f ms %.0f TFLOP/s" % (ms, fl / ms / 1e9)
    print(msg, "ok" if ok else "FAIL", flush=True)
    if not ok:
        e = (dw - ref).abs()
        per_tap = e.amax((1, 2)) / ref.abs().max()
        print("   per-tap err:", " ".join("%.0e" % v for v in per_tap.tolist()))
        per_ci = e.reshape(k * k, -1, 64, cout).amax((0, 2, 3)) / ref.abs().max()
        per_co = e.reshape(k * k, cin + cin2, -1, 64).amax((0, 1, 3)) / ref.abs().max()
        print("   per-ci-chunk:", per_ci.tolist(), "per-co-block:", per_co.tolist())
        t = int(per_tap.argmax())
        et = e[t]
        idx = (et == et.max()).nonzero()[0].tolist()
        print("   worst tap", t, "at (ci,co)", idx, "got", dw[t, idx[0], idx[1]].item(), "ref", ref[t, idx[0], idx[1]].item())
    return ok


def main():
    ok = True
    if "--k5" in sys.argv:
        run("wg halo 64->64 k5 16x16", 1, 16, 16, 64, 64, 5)
        run("wg halo 64->64 k5 32x40", 1, 32, 40, 64, 64, 5)
        run("wg halo 256->256 k5 32x40", 1, 32, 40, 256, 256, 5)
        run("wg halo 256->256 k5 32x48", 1, 32, 48, 256, 256, 5)
        run("wg halo 128->64 k5 32x48", 2, 32, 48, 128, 64, 5)
        return 0
    ok &= run("wg halo 64->64 k3 16x32", 2, 16, 32, 64, 64, 3)
    ok &= run("wg halo 64->64 k9 32x64", 2, 32, 64, 64, 64, 9)
    ok &= run("wg halo 128->128 k7 32x48", 2, 32, 48, 128, 128, 7)
    ok &= run("wg halo 256->256 k5 32x40", 1, 32, 40, 256, 256, 5)
    ok &= run("wg halo 512->512 k3 16x52", 3, 16, 52, 512, 512, 3)
    ok &= run("wg halo 512->512 k3 8x26", 5, 8, 26, 512, 512, 3)
    ok &= run("wg halo 128->64 k7 reflect", 2, 32, 64, 128, 64, 7, pad_mode="reflect")
    ok &= run("wg tap 64->128 k7 s2 reflect", 2, 32, 64, 64, 128, 7, stride=2, pad_mode="reflect")
    ok &= run("wg tap 256->512 k3 s2 zero", 2, 16, 40, 256, 512, 3, stride=2)
    ok &= run("wg tap 64->128 k4 s2 reflect", 2, 32, 64, 64, 128, 4, stride=2, pad_mode="reflect")
    ok &= run("wg tap 1x1 concat 128+128->128", 2, 16, 40, 128, 128, 1, cin2=128)
    ok &= run("wg tap 1x1 256->64", 2, 16, 40, 256, 64, 1)
    if "--big" in sys.argv:
        for nimg in (2, 8, 20):
            ok &= run("wg tap 1x1 concat 128+128->128 64x208 B%d" % nimg, nimg, 64, 208, 128, 128, 1, cin2=128)
        ok &= run("wg tap 1x1 concat 256+256->256 32x104 B20", 20, 32, 104, 256, 256, 1, cin2=256)
        ok &= run("wg tap 1x1 concat 512+512->512 16x52 B20", 20, 16, 52, 512, 512, 1, cin2=512)
        ok &= run("wg tap 256->512 k3 s2 32x104 B20", 20, 32, 104, 256, 512, 3, stride=2, pad_mode="reflect")
    if "--time" in sys.argv:
        run("T wg 64->64 k9 128x416 B20", 20, 128, 416, 64, 64, 9, timing=True)
        run("T wg 128->128 k7 64x208 B20", 20, 64, 208, 128, 128, 7, timing=True)
        run("T wg 256->256 k5 32x104 B20", 20, 32, 104, 256, 256, 5, timing=True)
        run("T wg 512->512 k3 16x52 B20", 20, 16, 52, 512, 512, 3, timing=True)
        run("T wg 512->512 k3 8x26 B20", 20, 8, 26, 512, 512, 3, timing=True)
        run("T wg 128->64 k7 128x416 B20", 20, 128, 416, 128, 64, 7, timing=True)
        run("T wg 64->128 k7 s2 B20", 20, 128, 416, 64, 128, 7, stride=2, pad_mode="reflect", timing=True)
        run("T wg 1x1 128+... 128->64 128x416", 20, 128, 416, 64, 64, 1, cin2=64, timing=True)
    print("ALL OK" if ok else "SOME FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
