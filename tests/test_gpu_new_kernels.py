"""GPU (-m gpu): the small kernels behind the guidance gradient, one by one against torch, and the CPU emulation of
the ABI (tests/abi_emulator.py) against the real kernels on whole engine plans -- the emulator is what the CPU suite
trusts for plan checks, so it is pinned here on the hardware it stands in for."""
import ctypes as C
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
dev = "cuda"


@pytest.mark.parametrize("relu,use_f32,scaled", [(1, True, True), (1, False, True), (0, True, True), (1, True, False)])
def test_act_backward_frozen(relu, use_f32, scaled):
    from gdn_pytorch_b200 import _lib
    from gdn_pytorch_b200.engine import FrozenBwdDesc
    g = torch.Generator().manual_seed(3)
    N, H, W, Cc = 3, 10, 14, 128
    dact = (torch.randn((N, H, W, Cc), generator=g) * 1e-4).to(dev)
    y = torch.relu(torch.randn((N, H, W, Cc), generator=g)).to(dev)
    yb = y.to(torch.bfloat16)
    scale = (torch.randn(Cc, generator=g)).to(dev)          # mixed signs on purpose
    dy = torch.full((N, H, W, Cc), float("nan"), device=dev, dtype=torch.bfloat16)
    d = FrozenBwdDesc()
    d.dact = dact.data_ptr()
    if relu:
        if use_f32:
            d.y_f32 = y.data_ptr()
        else:
            d.y_bf16 = yb.data_ptr()
    d.scale = scale.data_ptr() if scaled else None
    d.relu, d.n, d.h, d.w, d.c = relu, N, H, W, Cc
    d.dy = dy.data_ptr()
    _lib.check(_lib.lib().gdn_act_backward_frozen(C.byref(d), _lib.stream_ptr()), "frozen")
    ref = dact.clone()
    if relu:
        ref = ref * ((y if use_f32 else yb.float()) > 0)
    if scaled:
        ref = ref * scale
    assert torch.equal(dy, ref.to(torch.bfloat16))


@pytest.mark.parametrize("c,pad,reflect,acc", [(1, 4, 1, 0), (1, 4, 1, 1), (3, 2, 1, 0), (1, 3, 0, 0), (2, 0, 0, 1)])
def test_fold_grad_thin_channels(c, pad, reflect, acc):
    """adjoint of reflection / zero padding for 1..3-channel tensors (the input gradient of a first layer)"""
    from gdn_pytorch_b200 import _lib
    from gdn_pytorch_b200.engine import FoldDesc
    g = torch.Generator().manual_seed(c + pad)
    N, H, W = 2, 12, 20
    dpad = torch.randn((N, H + 2 * pad, W + 2 * pad, c), generator=g).to(dev)
    base = torch.randn((N, H, W, c), generator=g).to(dev)
    out = base.clone()
    f = FoldDesc()
    f.dpad, f.ctot, f.c_off = dpad.data_ptr(), c, 0
    f.n, f.h, f.w, f.c = N, H, W, c
    f.pad, f.reflect, f.up, f.dilate = pad, reflect, 0, 0
    f.dact, f.accumulate = out.data_ptr(), acc
    _lib.check(_lib.lib().gdn_fold_grad(C.byref(f), _lib.stream_ptr()), "fold")
    x = torch.zeros((N, c, H, W), device=dev, dtype=torch.float64, requires_grad=True)
    z = F.pad(x, (pad,) * 4, mode="reflect" if reflect else "constant") if pad else x
    (gx,) = torch.autograd.grad(z, x, dpad.double().permute(0, 3, 1, 2))
    ref = gx.permute(0, 2, 3, 1) + (base.double() if acc else 0)
    assert torch.allclose(out.double(), ref, rtol=1e-6, atol=1e-6)


def test_sqdiff_grad_and_tanh_chain_add():
    from gdn_pytorch_b200.ops import LossKernels
    g = torch.Generator().manual_seed(9)
    k = LossKernels(torch.device(dev))
    feats = [torch.randn((2, 8, 12, c), generator=g).to(dev) for c in (64, 128, 512, 512)]
    tars = [torch.randn(f.shape, generator=g).to(dev) for f in feats]
    grads = [torch.full_like(f, float("nan")) for f in feats]
    k.latent_grad(feats, tars, grads)
    fr = [f.clone().requires_grad_(True) for f in feats]
    lat = 1.5 * sum(w * F.mse_loss(a, b) for w, a, b in zip(k.LATENT_W, fr, tars)) / 4
    ref = torch.autograd.grad(lat, fr)
    for a, b in zip(grads, ref):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-12)
    out = torch.tanh(torch.randn((2, 1, 16, 24), generator=g)).to(dev)
    dout = torch.randn(out.shape, generator=g).to(dev)
    dpre = torch.randn(out.shape, generator=g).to(dev)
    want = dpre + 0.5 * dout * (1 - out * out)
    k.tanh_chain_add(dout, out, dpre, scale=0.5)
    assert torch.allclose(dpre, want, rtol=1e-6, atol=1e-7)


def _l2(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.mark.parametrize("gname,train", [("mini_rtod", False), ("mini_dtod", False), ("mini_deep512", False),
                                          ("mini_rtod", True), ("mini_dtod", True)])
def test_abi_emulator_agrees_with_the_kernels(gname, train):
    """same plan, same inputs: real kernels on the B200 vs tests/abi_emulator.py on the host.  Both sides round the
    operands to bf16 identically; differences are fp32 accumulation order (and, in training, the few ReLU decisions
    that order flips)."""
    from gdn_pytorch_b200.engine import Engine
    from oracle import synth
    from tests import minigraphs
    from tests.abi_emulator import emulated_abi, engine_forward, run_ops
    B, H, W = 2, 32, 64
    g = getattr(minigraphs, gname)()
    sd = minigraphs.synth_params(g, 0)
    x = synth.synth_rgb(B, H, W, 1) if g.cin == 3 else synth.synth_depth(B, H, W, 1)
    R = torch.rand((B, 1, H, W), generator=torch.Generator().manual_seed(5)) - 0.5
    names = [u.out for u in g.units]

    def params(device):
        p = {k: v.clone().to(device) for k, v in sd.items()}
        for k, v in p.items():
            if train and not k.endswith(("running_mean", "running_var")):
                v.requires_grad_(True)
        return p
    eng = Engine(g, params(dev), B, H, W, train=train, backward=train, want=names)
    with torch.no_grad():
        eng.forward(x.to(dev))
        if train:
            out = eng.depth()
            eng.flat_grad.zero_()
            eng.backward((R.to(dev) * (1 - out * out)).view(B, H, W))
    torch.cuda.synchronize()
    with emulated_abi():
        emu = Engine(g, params("cpu"), B, H, W, train=train, backward=train, want=names, device=torch.device("cpu"))
        with torch.no_grad():
            engine_forward(emu, x)
            if train:
                out = emu.depth()
                emu.flat_grad.zero_()
                emu.dpre.copy_((R * (1 - out * out)).view(B, H, W))
                run_ops(emu.bwd)
    # measured on the B200: eval <= 2e-3 on every hidden tensor, 2e-3..6e-3 on the tanh head at the end of the chain
    # (accumulation-order differences flip bf16 roundings downstream); training <= 3e-2 on activations.  Training
    # gradients are compared as ONE vector: a handful of ReLU decisions differ between the two runs (the comparison
    # against fp64 autograd imposes the engine's masks for exactly this reason; first-layer weights then differ by
    # ~0.1), and gradients that are analytically zero (BatchNorm shifts feeding another BatchNorm) are pure rounding
    # noise on both sides, so per-tensor relative errors are meaningless there.  A layout / offset bug gives O(1).
    tol = 3e-2 if train else 1e-2
    for n in names:
        assert _l2(eng.value(n).cpu(), emu.value(n)) <= tol, n
    if train:
        a, b = eng.flat_grad.cpu(), emu.flat_grad
        cos = (torch.dot(a, b) / (a.norm() * b.norm())).item()
        assert cos >= 0.97 and _l2(a, b) <= 0.25, (cos, _l2(a, b))


@pytest.mark.parametrize("cin,cout,k,stride2dst,resid,algo", [(64, 64, 1, False, False, 0), (128, 64, 1, False, True, 0),
                                                             (64, 128, 3, False, True, 2), (256, 256, 3, True, False, 1),
                                                             (64, 64, 9, False, True, 2 | (4 << 8) | (1 << 24))])
def test_fp32_only_epilogue_strided_destination(cin, cout, k, stride2dst, resid, algo):
    """fp32-only outputs with residual accumulation and strided destinations (what the input-gradient launches write)
    against an fp64 reference.  (A warp-transposed, line-per-pixel variant of this epilogue was measured on the B200 in
    round 2 -- 454 img/s against 550 for the whole step, profiles/r02a_* -- and removed.)"""
    import torch.nn.functional as F
    from gdn_pytorch_b200 import _lib
    g = torch.Generator().manual_seed(cin + cout + k)
    N, H, W = 2, 24, 40
    p = k // 2
    x = (torch.rand((N, cin, H, W), generator=g) * 2 - 1).to(dev).to(torch.bfloat16)
    w = ((torch.rand((cout, cin, k, k), generator=g) * 2 - 1) / (cin * k * k) ** 0.5).to(dev).to(torch.bfloat16)
    ref = F.conv2d(x.double(), w.double(), None, 1, p)
    xb = x.permute(0, 2, 3, 1).contiguous()
    wp = w.permute(2, 3, 0, 1).reshape(k * k, cout, cin).contiguous()
    sy = 2 if stride2dst else 1
    DH, DW = H * sy, W * sy
    out = torch.full((N, DH, DW, cout), float("nan"), device=dev)
    r = (torch.rand((N, DH, DW, cout), generator=g) - 0.5).to(dev) if resid else None
    d = _lib.ConvDesc()
    d.src0 = _lib.Act(xb.data_ptr(), N, H, W, cin, 0)
    d.weights = wp.data_ptr()
    d.kh = d.kw = k
    d.stride = 1
    d.off_y = d.off_x = -p
    d.out_h, d.out_w, d.cout, d.cout_pad, d.algo = H, W, cout, cout, algo
    d.resid = r.data_ptr() if resid else None
    d.out_f32 = out.data_ptr()
    d.dst_h, d.dst_w, d.dst_sy, d.dst_sx, d.dst_oy, d.dst_ox = DH, DW, sy, sy, sy - 1, 0
    _lib.check(_lib.lib().gdn_conv2d(C.byref(d), _lib.stream_ptr()), "conv")
    torch.cuda.synchronize()
    got = out[:, sy - 1::sy, ::sy].permute(0, 3, 1, 2).double()
    want = ref + (r[:, sy - 1::sy, ::sy].permute(0, 3, 1, 2).double() if resid else 0)
    assert not torch.isnan(got).any()
    assert (got - want).abs().max().item() <= 1e-3 * ref.abs().max().item()
    if stride2dst:                                   # positions between the strided destinations stay untouched
        assert torch.isnan(out[:, 0::2, 1::2]).all()


@pytest.mark.parametrize("cin,cout,k,relu,resid,keep32,algo,N,H,W", [
    (64, 64, 3, True, False, False, 0, 2, 24, 40),
    (128, 64, 1, True, True, False, 0, 2, 24, 40),
    (64, 128, 3, False, True, True, 2 | (2 << 8), 3, 16, 24),
    (512, 512, 3, True, False, False, 1 | (1 << 24), 4, 8, 26),
    (64, 64, 9, True, True, False, 2 | (4 << 8) | (1 << 24), 4, 64, 96),
    (512, 512, 3, True, True, True, 1 | (1 << 24) | (2 << 25), 4, 8, 26),         # split-K: the combine kernel does it
    (512, 512, 3, False, False, False, 1 | (4 << 25), 3, 8, 26)])
def test_fused_bn_backward_statistics_epilogue(cin, cout, k, relu, resid, keep32, algo, N, H, W):
    """gdn_conv_desc.bwd_raw: the input-gradient launch that completes a tensor's gradient applies the producer's ReLU mask
    and reduces sum g / sum g*xhat in its epilogue.  Against the separate path: fp64 conv (+ accumulate), mask from the
    same fma expression, sums in fp64; outputs are the masked gradient as bf16 (and fp32 when the identity branch needs it)"""
    import torch.nn.functional as F
    from gdn_pytorch_b200 import _lib
    g = torch.Generator().manual_seed(cin + cout + k + H)
    p = k // 2
    dy = (torch.rand((N, cin, H, W), generator=g) * 2 - 1).to(dev).to(torch.bfloat16)
    w = ((torch.rand((cout, cin, k, k), generator=g) * 2 - 1) / (cin * k * k) ** 0.5).to(dev).to(torch.bfloat16)
    tot = F.conv2d(dy.double(), w.double(), None, 1, p).permute(0, 2, 3, 1)                 # NHWC fp64
    r = (torch.rand((N, H, W, cout), generator=g) - 0.5).to(dev) if resid else None
    if resid:
        tot = tot + r.double()
    raw = (torch.randn((N, H, W, cout), generator=g) * 1.5 + 0.2).to(dev).half()
    coef = torch.stack([torch.rand(cout, generator=g) + 0.5, torch.rand(cout, generator=g) - 0.5,
                        torch.rand(cout, generator=g) * 0.4, torch.rand(cout, generator=g) + 0.5], 1).to(dev).contiguous()
    rawf = raw.float()
    y64 = rawf.double() * coef[:, 0].double() + coef[:, 1].double()
    mask = (y64 > 0) if relu else torch.ones_like(rawf, dtype=torch.bool)
    safe = (y64.abs() > 1e-5) if relu else mask          # a mask test at fp32 rounding level may legitimately flip
    xhat = ((rawf - coef[:, 2]) * coef[:, 3]).double()
    gmask = tot * mask
    want_s1, want_s2 = gmask.sum((0, 1, 2)), (gmask * xhat).sum((0, 1, 2))
    sums = torch.zeros((2, cout), dtype=torch.float64, device=dev)
    out16 = torch.full((N, H, W, cout), float("nan"), device=dev, dtype=torch.bfloat16)
    out32 = torch.full((N, H, W, cout), float("nan"), device=dev) if keep32 else None
    d = _lib.ConvDesc()
    xb = dy.permute(0, 2, 3, 1).contiguous()
    d.src0 = _lib.Act(xb.data_ptr(), N, H, W, cin, 0)
    wp = w.permute(2, 3, 0, 1).reshape(k * k, cout, cin).contiguous()
    d.weights = wp.data_ptr()
    d.kh = d.kw = k
    d.stride = 1
    d.off_y = d.off_x = -p
    d.out_h, d.out_w, d.cout, d.cout_pad, d.algo = H, W, cout, cout, algo
    d.dst_h, d.dst_w, d.dst_sy, d.dst_sx = H, W, 1, 1
    d.resid = r.data_ptr() if resid else None
    d.out_bf16 = _lib.Act(out16.data_ptr(), N, H, W, cout, 0)
    d.out_f32 = out32.data_ptr() if keep32 else None
    d.bwd_raw, d.bwd_coef, d.bwd_relu = raw.data_ptr(), coef.data_ptr(), int(relu)
    d.stat_sum, d.stat_sqsum = sums[0].data_ptr(), sums[1].data_ptr()
    if (algo >> 25) & 7:
        ws = torch.empty(_lib.lib().gdn_conv2d_workspace_bytes(C.byref(d)), dtype=torch.uint8, device=dev)
        d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
    _lib.check(_lib.lib().gdn_conv2d(C.byref(d), _lib.stream_ptr()), "conv")
    torch.cuda.synchronize()
    scale_ = tot.abs().max().item()
    # elements whose mask test sits at rounding level cannot flip here (the mask comes from the same stored fp16 / fp32 values)
    assert ((out16.double() - gmask) * safe).abs().max().item() <= (2 ** -8) * scale_
    if keep32:
        assert ((out32.double() - gmask) * safe).abs().max().item() <= 1e-3 * scale_
    for got, want in ((sums[0], want_s1), (sums[1], want_s2)):
        assert (got - want).abs().max().item() <= 2e-3 * want.abs().max().item() + 1e-6, (got - want).abs().max().item()
    d.bwd_coef = None
    assert _lib.lib().gdn_conv2d(C.byref(d), _lib.stream_ptr()) == -1            # incomplete descriptor is refused


@pytest.mark.parametrize("n,h,w,half,tanh", [(2, 13, 45, 1, 1), (1, 8, 32, 1, 0), (3, 16, 64, 0, 1), (2, 128, 416, 1, 1)])
def test_head_gather_matches_shifted_sum(n, h, w, half, tanh):
    """gdn_head_gather: out[p] = act(sum_t Z[p + off(t)][t]) with out-of-image taps skipped -- ragged extents (tiles are
    8 x 32), both 16-bit formats, and the bench extent; same fp32 summation order on both sides -> exact"""
    from gdn_pytorch_b200 import _lib
    g = torch.Generator().manual_seed(n * 1000 + h)
    k, pad, zc = 9, 4, 128
    z = torch.randn((n, h, w, zc), generator=g).to(dev).to(torch.float16 if half else torch.bfloat16)
    out = torch.full((n, h, w), float("nan"), device=dev)
    _lib.check(_lib.lib().gdn_head_gather(C.c_void_p(z.data_ptr()), half, zc, n, h, w, k, pad, tanh,
                                          C.c_void_p(out.data_ptr()), _lib.stream_ptr()), "head_gather")
    zp = F.pad(z.float().permute(0, 3, 1, 2), (pad, k - 1 - pad, pad, k - 1 - pad))
    acc = torch.zeros((n, h, w), device=dev)
    for r in range(k):
        for s in range(k):
            acc += zp[:, r * k + s, r:r + h, s:s + w]
    if tanh:
        assert (out - torch.tanh(acc)).abs().max().item() <= 2e-6       # tanhf vs torch.tanh: a few ulp
    else:
        assert torch.equal(out, acc)


@pytest.mark.parametrize("transposed", [False, True])
def test_head_as_taps_matches_conv2d(transposed):
    """the whole head (pack with the taps as output channels -> 1x1 gdn_conv2d -> gdn_head_gather), through a one-unit
    engine graph, against F.conv2d / F.conv_transpose2d on the bf16-rounded operands (AE_model_unet.py:300 / :521)"""
    from gdn_pytorch_b200.engine import Engine
    from gdn_pytorch_b200.graph import Graph, Unit
    g = torch.Generator().manual_seed(11)
    N, H, W = 2, 24, 40
    x = torch.randn((N, 64, H, W), generator=g).to(dev)
    wt = (torch.randn((64, 1, 9, 9) if transposed else (1, 64, 9, 9), generator=g) * 0.02).to(dev)
    gr = Graph("head_only", 64, [Unit("head", ("in",), "y", 64, 1, 9, 1, 4, transposed=transposed, tanh=True)], ("y",))
    eng = Engine(gr, {"head.weight": wt}, N, H, W, train=False, want=("y",))
    assert hasattr(eng.cu["head"], "z"), "the head did not take the taps-as-N path"
    eng.forward(x)
    got = eng.value("y").reshape(N, H, W)
    xb, wb = x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float()
    ref = torch.tanh(F.conv_transpose2d(xb, wb, padding=4) if transposed else F.conv2d(xb, wb, padding=4)).reshape(N, H, W)
    # fp16 storage of the 81 per-tap partial sums: 2^-11 relative each
    assert (got - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())
