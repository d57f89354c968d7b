"""GPU check of gdn_conv2d against torch fp32 (run on a B200 box).  Usage: python tools/check_conv.py"""
import ctypes as C
import sys, os, time
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


def pack_w(w):  # OIHW fp32 -> [kh*kw][O][I] bf16
    o, i, kh, kw = w.shape
    return w.permute(2, 3, 0, 1).reshape(kh * kw, o, i).contiguous().to(torch.bfloat16)


def run_case(name, N, H, W, cin, cout, k, stride=1, pad_mode="zero", algo=0, relu=False, bias=False, resid=False,
             stats=False, reflect_out=0, cin2=0, timing=False):
    g = torch.Generator(device="cpu").manual_seed(hash(name) % 1000)
    p = k // 2
    x = (torch.rand((N, cin + cin2, H, W), generator=g) * 2 - 1).to(dev)
    w = ((torch.rand((cout, cin + cin2, k, k), generator=g) * 2 - 1) / (cin * k * k) ** 0.5).to(dev)
    xb = x.to(torch.bfloat16).float()
    wb = w.to(torch.bfloat16).float()
    if pad_mode == "reflect":
        xin = F.pad(xb, (p,) * 4, mode="reflect")
        ref = F.conv2d(xin, wb, None, stride, 0)
        bufpad = p
        xbuf = xin.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        off = -p
    else:
        ref = F.conv2d(xb, wb, None, stride, p)
        bufpad = 0
        xbuf = xb.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        off = -p
    OH, OW = ref.shape[2], ref.shape[3]
    raw = ref.clone()
    b = None
    if bias:
        b = (torch.rand(cout, generator=g) - 0.5).to(dev)
        ref = ref + b.view(1, -1, 1, 1)
    if relu:
        ref = F.relu(ref)
    r = None
    if resid:
        r = (torch.rand((N, OH, OW, cout), generator=g) - 0.5).to(dev)
        ref = ref + r.permute(0, 3, 1, 2)
    d = _lib.ConvDesc()
    if cin2:
        x0 = xbuf[..., :cin].contiguous()
        x1 = xbuf[..., cin:].contiguous()
        d.src0 = act_of(x0, bufpad)
        d.src1 = act_of(x1, bufpad)
    else:
        d.src0 = act_of(xbuf, bufpad)
    wp = pack_w(wb)
    d.weights = wp.data_ptr()
    d.kh = d.kw = k
    d.stride = stride
    d.off_y = d.off_x = off
    d.out_h, d.out_w = OH, OW
    d.cout = cout
    d.cout_pad = max(cout, 16)
    if cout < 16:
        wpad = torch.zeros((k * k, 16, cin + cin2), dtype=torch.bfloat16, device=dev)
        wpad[:, :cout] = wp
        wp = wpad
        d.weights = wp.data_ptr()
    d.algo = algo
    d.bias = b.data_ptr() if bias else None
    d.relu = int(relu)
    d.resid = r.data_ptr() if resid else None
    out32 = torch.full((N, OH, OW, cout), float("nan"), device=dev)
    d.out_f32 = out32.data_ptr()
    P = reflect_out
    outb = torch.full((N, OH + 2 * P, OW + 2 * P, cout), float("nan"), device=dev, dtype=torch.bfloat16)
    d.out_bf16 = act_of(outb, P)
    d.out_reflect = 1 if P else 0
    d.dst_h, d.dst_w = OH, OW
    d.dst_sy = d.dst_sx = 1
    ssum = torch.zeros(cout, dtype=torch.float64, device=dev)
    ssq = torch.zeros(cout, dtype=torch.float64, device=dev)
    if stats:
        d.stat_sum, d.stat_sqsum = ssum.data_ptr(), ssq.data_ptr()
    L = _lib.lib()
    rc = L.gdn_conv2d(C.byref(d), _lib.stream_ptr())
    _lib.check(rc, name)
    torch.cuda.synchronize()
    got = out32.permute(0, 3, 1, 2)
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item() / scale
    ok = err < 2e-3 and not torch.isnan(got).any().item()
    msg = "%-34s err=%.2e" % (name, err)
    gb = outb.float()
    if P:
        refp = F.pad(ref, (P,) * 4, mode="reflect").permute(0, 2, 3, 1)
    else:
        refp = ref.permute(0, 2, 3, 1)
    errb = (gb - refp).abs().max().item() / scale
    ok = ok and errb < 1e-2 and not torch.isnan(gb).any().item()
    msg += " bf16err=%.2e" % errb
    if stats:
        rs = raw.double().sum((0, 2, 3))
        rq = (raw.double() ** 2).sum((0, 2, 3))
        e1 = ((ssum - rs).abs().max() / rs.abs().max()).item()
        e2 = ((ssq - rq).abs().max() / rq.abs().max()).item()
        ok = ok and e1 < 1e-3 and e2 < 1e-3
        msg += " stat=%.1e/%.1e" % (e1, e2)
    if timing:
        for _ in range(3):
            L.gdn_conv2d(C.byref(d), _lib.stream_ptr())
        torch.cuda.synchronize()
        e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            L.gdn_conv2d(C.byref(d), _lib.stream_ptr())
        e1_.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1_) / reps
        fl = 2.0 * N * OH * OW * cout * (cin + cin2) * k * k
        msg += "  %.3f ms %.0f TFLOP/s" % (ms, fl / ms / 1e9)
    print(msg, "ok" if ok else "FAIL", flush=True)
    return ok


def main():
    ok = True
    HALO, TAP = 2, 1
    # small correctness cases
    ok &= run_case("tap 64->64 k3 16x24", 2, 16, 24, 64, 64, 3, algo=TAP)
    ok &= run_case("halo 64->64 k3 16x32", 2, 16, 32, 64, 64, 3, algo=HALO)
    ok &= run_case("halo 64->64 k9 32x64", 2, 32, 64, 64, 64, 9, algo=HALO, stats=True)
    ok &= run_case("halo 128->128 k7 32x48", 2, 32, 48, 128, 128, 7, algo=HALO, relu=True, bias=True)
    ok &= run_case("halo 256->256 k5 32x40 resid", 1, 32, 40, 256, 256, 5, algo=HALO, resid=True, stats=True)
    ok &= run_case("halo 512->512 k3 16x52", 3, 16, 52, 512, 512, 3, algo=HALO, stats=True)
    ok &= run_case("tap 512->512 k3 8x26", 5, 8, 26, 512, 512, 3, algo=TAP, relu=True, bias=True, resid=True, stats=True)
    ok &= run_case("tap 64->128 k7 s2 reflect", 2, 32, 64, 64, 128, 7, stride=2, pad_mode="reflect", algo=TAP)
    ok &= run_case("tap 256->512 k3 s2 zero", 2, 16, 40, 256, 512, 3, stride=2, algo=TAP)
    ok &= run_case("tap 1x1 concat 128+128->128", 2, 16, 40, 128, 128, 1, algo=TAP, cin2=128, stats=True)
    ok &= run_case("halo 128->64 k7 reflect in", 2, 32, 64, 128, 64, 7, pad_mode="reflect", algo=HALO, reflect_out=3)
    ok &= run_case("halo 64->1 k9 head", 2, 32, 64, 64, 1, 9, algo=HALO)
    ok &= run_case("tap 64->64 k4 s2 reflect1", 2, 32, 64, 64, 128, 4, stride=2, pad_mode="reflect", algo=TAP) if False else True
    ok &= run_case("auto 64->64 k9 reflect_out4", 1, 48, 72, 64, 64, 9, reflect_out=4, relu=True)
    # full-size timing of the hot shapes (B=20)
    if "--time" in sys.argv:
        run_case("T halo 64->64 k9 128x416 B20", 20, 128, 416, 64, 64, 9, algo=HALO, stats=True, timing=True)
        run_case("T tap  64->64 k9 128x416 B20", 20, 128, 416, 64, 64, 9, algo=TAP, timing=True)
        run_case("T halo 128->128 k7 64x208 B20", 20, 64, 208, 128, 128, 7, algo=HALO, stats=True, timing=True)
        run_case("T halo 256->256 k5 32x104 B20", 20, 32, 104, 256, 256, 5, algo=HALO, stats=True, timing=True)
        run_case("T halo 512->512 k3 16x52 B20", 20, 16, 52, 512, 512, 3, algo=HALO, timing=True)
        run_case("T tap  512->512 k3 16x52 B20", 20, 16, 52, 512, 512, 3, algo=TAP, timing=True)
        run_case("T tap  512->512 k3 8x26 B20", 20, 8, 26, 512, 512, 3, algo=TAP, timing=True)
        run_case("T halo 128->64 k7 128x416 B20", 20, 128, 416, 128, 64, 7, algo=HALO, timing=True)
        run_case("T halo 64->1 k9 128x416 B20", 20, 128, 416, 64, 1, 9, algo=HALO, timing=True)
    print("ALL OK" if ok else "SOME FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
