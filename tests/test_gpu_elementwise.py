"""GPU (-m gpu): the HBM-bound helper kernels (gdn_act_forward, gdn_bn_bwd_reduce / gdn_act_backward, gdn_fold_grad,
gdn_im2col, gdn_pack_weights / gdn_unpack_wgrad) through the C ABI against plain torch fp32/fp64 references of the
ops they replace (BatchNorm apply + ReLU + residual + ReflectionPad2d + F.interpolate and their autograd;
/root/reference/src/AE_model_unet.py:45-94, 336-355).  Tolerances: bf16 storage = 2^-8 relative, fp32 paths 1e-5."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
dev = "cuda"


def _L():
    from gdn_pytorch_b200 import _lib
    return _lib, _lib.lib()


def _transform(y, pad, reflect, up, dilate):
    """the input transform a consuming conv sees: (N,C,H,W) -> padded / upsampled / dilated (N,C,H',W')"""
    if up:
        y = F.interpolate(y, scale_factor=2, mode="bilinear", align_corners=(up == 2))
    if dilate:
        z = torch.zeros((y.shape[0], y.shape[1], 2 * y.shape[2], 2 * y.shape[3]), dtype=y.dtype, device=y.device)
        z[:, :, ::2, ::2] = y
        y = z
    if pad:
        y = F.pad(y, (pad,) * 4, mode="reflect") if reflect else F.pad(y, (pad,) * 4)
    return y


ACT_CASES = [
    # n, h, w, c, src, scale, relu, resid, f32, b16, pad, reflect, up, dilate
    (2, 16, 24, 64, "half", True, True, False, False, True, 0, 0, 0, 0),
    (2, 16, 24, 64, "half", True, False, True, True, True, 0, 0, 0, 0),
    (3, 10, 13, 128, "half", True, True, False, True, True, 3, 1, 0, 0),
    (2, 9, 20, 64, "f32", False, False, False, False, True, 4, 1, 0, 0),
    (2, 8, 26, 512, "half", True, True, False, False, True, 1, 1, 1, 0),
    (2, 12, 10, 256, "half", True, False, True, True, True, 2, 1, 1, 0),
    (1, 6, 7, 64, "bf16", True, True, False, False, True, 0, 0, 1, 0),
    (2, 8, 12, 128, "half", True, True, False, False, True, 0, 0, 0, 1),
    (2, 8, 12, 64, "f32", False, False, False, False, True, 3, 1, 2, 0),
    (20, 128, 416, 64, "half", True, True, False, False, True, 0, 0, 0, 0),
    (4, 64, 208, 128, "half", True, False, True, True, True, 3, 1, 1, 0),
    # pure x2 bilinear of a plain bf16 tensor (+ reflection border): the source-block-driven kernel (up2x_blocks_kernel)
    (2, 8, 26, 512, "bf16", False, False, False, False, True, 1, 1, 1, 0),
    (3, 9, 7, 64, "bf16", False, False, False, False, True, 3, 1, 1, 0),
    (2, 6, 10, 128, "bf16", False, False, False, False, True, 0, 0, 1, 0),
    (2, 5, 6, 64, "bf16", False, False, False, False, True, 2, 0, 1, 0),
    (20, 64, 208, 128, "bf16", False, False, False, False, True, 3, 1, 1, 0),
]


@pytest.mark.parametrize("case", ACT_CASES, ids=lambda c: "-".join(str(v) for v in c))
def test_act_forward(case):
    from gdn_pytorch_b200.engine import ActFwdDesc
    _lib, L = _L()
    n, h, w, c, src, scale, relu, resid, f32, b16, pad, reflect, up, dilate = case
    g = torch.Generator().manual_seed(1)
    x = torch.randn((n, h, w, c), generator=g).to(dev)
    a = ActFwdDesc()
    if src == "half":
        xs = x.half()
        a.src_bf16, a.src16_is_half = xs.data_ptr(), 1
    elif src == "bf16":
        xs = x.bfloat16()
        a.src_bf16 = xs.data_ptr()
    else:
        xs = x
        a.src_f32 = xs.data_ptr()
    ref = xs.double()
    if scale:
        sc = (torch.rand(c, generator=g) + 0.5).to(dev)
        sh = (torch.rand(c, generator=g) - 0.5).to(dev)
        a.scale, a.shift = sc.data_ptr(), sh.data_ptr()
        ref = ref * sc.double() + sh.double()
    if relu:
        ref = F.relu(ref)
        a.relu = 1
    if resid:
        r = torch.randn((n, h, w, c), generator=g).to(dev)
        a.resid = r.data_ptr()
        ref = ref + r.double()
    a.n, a.h, a.w, a.c = n, h, w, c
    a.pad, a.reflect, a.up, a.dilate = pad, reflect, up, dilate
    s = 2 if (up or dilate) else 1
    o32 = torch.full((n, h, w, c), float("nan"), device=dev)
    o16 = torch.full((n, h * s + 2 * pad, w * s + 2 * pad, c), float("nan"), device=dev, dtype=torch.bfloat16)
    if f32:
        a.out_f32 = o32.data_ptr()
    if b16:
        a.out_bf16 = o16.data_ptr()
    _lib.check(L.gdn_act_forward(C.byref(a), _lib.stream_ptr()), "act_forward")
    torch.cuda.synchronize()
    scale_ = ref.abs().max().item()
    if f32:
        assert (o32.double() - ref).abs().max().item() <= 1e-5 * scale_
    if b16:
        want = _transform(ref.permute(0, 3, 1, 2), pad, reflect, up, dilate).permute(0, 2, 3, 1)
        got = o16.double()
        if pad and not reflect:
            got, want = got[:, pad:-pad, pad:-pad], want[:, pad:-pad, pad:-pad]
        assert not torch.isnan(got).any()
        assert (got - want).abs().max().item() <= 5e-3 * scale_


@pytest.mark.parametrize("shape,relu,half", [((2, 16, 24, 64), True, True), ((3, 9, 13, 128), False, True),
                                              ((2, 8, 26, 512), True, False), ((20, 64, 208, 128), True, True)])
def test_bn_backward_reduce_and_apply(shape, relu, half):
    """BatchNorm(+ReLU) backward against autograd of F.batch_norm(training=True) in fp64"""
    from gdn_pytorch_b200.engine import BnBwdDesc
    _lib, L = _L()
    n, h, w, c = shape
    g = torch.Generator().manual_seed(2)
    raw = torch.randn((n, h, w, c), generator=g).to(dev) * 1.5 + 0.3
    raw16 = raw.half() if half else raw.bfloat16()
    gamma = (torch.rand(c, generator=g) + 0.5).to(dev)
    beta = (torch.rand(c, generator=g) - 0.5).to(dev)
    dact = torch.randn((n, h, w, c), generator=g).to(dev)
    x = raw16.double().permute(0, 3, 1, 2).requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    y = F.batch_norm(x, None, None, gd, bd, True, 0.1, 1e-5)
    if relu:
        y = F.relu(y)
    y.backward(dact.double().permute(0, 3, 1, 2))
    mean = raw16.double().mean((0, 1, 2))
    var = raw16.double().var((0, 1, 2), unbiased=False)
    rstd = 1.0 / torch.sqrt(var + 1e-5)
    scale = (gamma.double() * rstd).float()
    shift = (beta.double() - mean * gamma.double() * rstd).float()
    meanf, rstdf = mean.float(), rstd.float()
    sums = torch.zeros((2, c), dtype=torch.float64, device=dev)
    dy = torch.full((n, h, w, c), float("nan"), device=dev, dtype=torch.bfloat16)
    dgamma, dbeta = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    b = BnBwdDesc()
    b.dact, b.raw = dact.data_ptr(), raw16.data_ptr()
    b.scale, b.shift, b.mean, b.rstd = scale.data_ptr(), shift.data_ptr(), meanf.data_ptr(), rstdf.data_ptr()
    b.relu, b.raw_is_half = int(relu), int(half)
    b.n, b.h, b.w, b.c = n, h, w, c
    b.sum_g, b.sum_gx = sums[0].data_ptr(), sums[1].data_ptr()
    b.dy = dy.data_ptr()
    b.dgamma, b.dbeta = dgamma.data_ptr(), dbeta.data_ptr()
    _lib.check(L.gdn_bn_bwd_reduce(C.byref(b), _lib.stream_ptr()), "bn_bwd_reduce")
    _lib.check(L.gdn_act_backward(C.byref(b), _lib.stream_ptr()), "act_backward")
    torch.cuda.synchronize()
    # ReLU masks can flip where |bn(x)| is at fp32 rounding level; compare in L2
    def l2(a_, b_):
        return ((a_ - b_).norm() / (b_.norm() + 1e-30)).item()
    assert l2(dbeta.double(), bd.grad) <= 1e-3
    assert l2(dgamma.double(), gd.grad) <= 1e-3
    assert l2(dy.double().permute(0, 3, 1, 2), x.grad) <= 6e-3
    # apply pass from a PRE-MASKED bf16 gradient with the sums already reduced (the fused-statistics plan): same dy
    mask = ((raw16.float() * scale + shift) > 0) if relu else torch.ones_like(dact, dtype=torch.bool)
    gm = (dact * mask).bfloat16()
    dy2 = torch.full((n, h, w, c), float("nan"), device=dev, dtype=torch.bfloat16)
    b.dact, b.dact_is_bf16, b.relu, b.dy = gm.data_ptr(), 1, 0, dy2.data_ptr()
    b.dgamma = b.dbeta = None
    _lib.check(L.gdn_act_backward(C.byref(b), _lib.stream_ptr()), "act_backward bf16")
    torch.cuda.synchronize()
    assert l2(dy2.double().permute(0, 3, 1, 2), x.grad) <= 8e-3


FOLD_CASES = [
    # n, h, w, c, ctot, c_off, pad, reflect, up, dilate, accumulate
    (2, 16, 24, 64, 64, 0, 3, 1, 0, 0, 0),
    (2, 16, 24, 64, 128, 64, 1, 1, 0, 0, 1),
    (2, 8, 26, 128, 128, 0, 1, 1, 1, 0, 0),
    (2, 9, 7, 64, 64, 0, 3, 1, 1, 0, 1),
    (2, 8, 12, 64, 64, 0, 0, 0, 0, 1, 0),
    (2, 6, 10, 256, 256, 0, 0, 0, 1, 0, 0),
    (4, 64, 208, 128, 128, 0, 3, 1, 1, 0, 0),
]


@pytest.mark.parametrize("case", FOLD_CASES, ids=lambda c: "-".join(str(v) for v in c))
def test_fold_grad_is_the_adjoint_of_the_input_transform(case):
    from gdn_pytorch_b200.engine import FoldDesc
    _lib, L = _L()
    n, h, w, c, ctot, c_off, pad, reflect, up, dilate, acc = case
    g = torch.Generator().manual_seed(3)
    s = 2 if (up or dilate) else 1
    dpad = torch.randn((n, h * s + 2 * pad, w * s + 2 * pad, ctot), generator=g).to(dev)
    prev = torch.randn((n, h, w, c), generator=g).to(dev)
    x = torch.zeros((n, c, h, w), dtype=torch.float64, device=dev, requires_grad=True)
    t = _transform(x, pad, reflect, up, dilate)
    t.backward(dpad[..., c_off:c_off + c].double().permute(0, 3, 1, 2))
    want = x.grad.permute(0, 2, 3, 1) + (prev.double() if acc else 0)
    dact = prev.clone()
    f = FoldDesc()
    f.dpad, f.ctot, f.c_off = dpad.data_ptr(), ctot, c_off
    f.n, f.h, f.w, f.c = n, h, w, c
    f.pad, f.reflect, f.up, f.dilate = pad, reflect, up, dilate
    f.dact, f.accumulate = dact.data_ptr(), acc
    _lib.check(L.gdn_fold_grad(C.byref(f), _lib.stream_ptr()), "fold")
    torch.cuda.synchronize()
    assert (dact.double() - want).abs().max().item() <= 1e-5 * want.abs().max().item()
    # the same adjoint from a bf16 operand gradient (what the input-gradient convolutions write): identical arithmetic on
    # the bf16-rounded values
    d16 = dpad.bfloat16()
    x2 = torch.zeros((n, c, h, w), dtype=torch.float64, device=dev, requires_grad=True)
    _transform(x2, pad, reflect, up, dilate).backward(d16[..., c_off:c_off + c].double().permute(0, 3, 1, 2))
    want2 = x2.grad.permute(0, 2, 3, 1) + (prev.double() if acc else 0)
    dact2 = prev.clone()
    f.dpad, f.dpad_is_bf16, f.dact = d16.data_ptr(), 1, dact2.data_ptr()
    _lib.check(L.gdn_fold_grad(C.byref(f), _lib.stream_ptr()), "fold bf16")
    torch.cuda.synchronize()
    assert (dact2.double() - want2).abs().max().item() <= 1e-5 * want2.abs().max().item()


@pytest.mark.parametrize("n,c,h,w,k,pad,reflect,kpad", [(2, 3, 16, 24, 9, 4, 1, 256), (2, 1, 12, 20, 9, 4, 1, 128),
                                                        (3, 1, 16, 16, 9, 4, 0, 128), (20, 3, 128, 416, 9, 4, 1, 256)])
def test_im2col(n, c, h, w, k, pad, reflect, kpad):
    _lib, L = _L()
    g = torch.Generator().manual_seed(4)
    x = torch.randn((n, c, h, w), generator=g).to(dev)
    col = torch.full((n, h, w, kpad), float("nan"), device=dev, dtype=torch.bfloat16)
    _lib.check(L.gdn_im2col(C.c_void_p(x.data_ptr()), C.c_void_p(col.data_ptr()), n, c, h, w, k, k, pad, reflect, kpad,
                            _lib.stream_ptr()), "im2col")
    torch.cuda.synchronize()
    xp = F.pad(x, (pad,) * 4, mode="reflect") if reflect else F.pad(x, (pad,) * 4)
    u = F.unfold(xp, k).view(n, c, k * k, h, w)                    # [n][c][tap][y][x]
    want = u.permute(0, 3, 4, 2, 1).reshape(n, h, w, k * k * c)     # column = tap*c + ch
    assert torch.equal(col[..., :k * k * c], want.bfloat16())
    assert (col[..., k * k * c:] == 0).all()


@pytest.mark.parametrize("cout,cin,k,transposed", [(64, 64, 9, False), (128, 64, 7, False), (512, 256, 3, False),
                                                   (128, 256, 1, False), (256, 512, 4, True), (64, 128, 5, True)])
def test_pack_weights_and_unpack_wgrad(cout, cin, k, transposed):
    """forward pack, dgrad pack (both channel halves of a virtual concat) and the wgrad scatter against torch permutes"""
    from gdn_pytorch_b200.engine import PackDesc
    _lib, L = _L()
    kk = k * k
    g = torch.Generator().manual_seed(5)
    if transposed:
        w = torch.randn((cin, cout, k, k), generator=g).to(dev)       # ConvTranspose2d: (cin, cout, k, k)
        pd = PackDesc(k, k, cout, cin, cout, cin, kk, cout * kk, k, 1, 1, 0)
        want = w.flip(2, 3).permute(2, 3, 1, 0).reshape(kk, cout, cin)
    else:
        w = torch.randn((cout, cin, k, k), generator=g).to(dev)
        pd = PackDesc(k, k, cout, cin, cout, cin, cin * kk, kk, k, 1, 0, 0)
        want = w.permute(2, 3, 0, 1).reshape(kk, cout, cin)
    scale = (torch.rand(cout, generator=g) + 0.5).to(dev)
    for sc in (None, scale):
        out = torch.full((kk, cout, cin), float("nan"), device=dev, dtype=torch.bfloat16)
        _lib.check(L.gdn_pack_weights(C.byref(pd), C.c_void_p(w.data_ptr()), C.c_void_p(sc.data_ptr() if sc is not None else None),
                                      C.c_void_p(out.data_ptr()), _lib.stream_ptr()), "pack")
        torch.cuda.synchronize()
        ref = want if sc is None else want * sc.view(1, -1, 1)
        assert torch.equal(out, ref.bfloat16())
    # dgrad pack of the second half of the input channels: [tap][ci][co]
    half = cin // 2
    if transposed:
        pdg = PackDesc(k, k, half, cout, half, cout, cout * kk, kk, k, 1, 0, 0)
        w_off = half * cout * kk
        wantd = w[half:].permute(2, 3, 0, 1).reshape(kk, half, cout)
    else:
        pdg = PackDesc(k, k, half, cout, half, cout, kk, cin * kk, k, 1, 1, 0)
        w_off = half * kk
        wantd = w[:, half:].flip(2, 3).permute(2, 3, 1, 0).reshape(kk, half, cout)
    outd = torch.full((kk, half, cout), float("nan"), device=dev, dtype=torch.bfloat16)
    _lib.check(L.gdn_pack_weights(C.byref(pdg), C.c_void_p(w.data_ptr() + 4 * w_off), C.c_void_p(None),
                                  C.c_void_p(outd.data_ptr()), _lib.stream_ptr()), "pack dgrad")
    torch.cuda.synchronize()
    assert torch.equal(outd, wantd.bfloat16())
    # wgrad scatter: dw[tap][ci][co] -> parameter layout (accumulating)
    dw = torch.randn((kk, cin, cout), generator=g).to(dev)
    grad = torch.randn(w.shape, generator=g).to(dev)
    before = grad.clone()
    _lib.check(L.gdn_unpack_wgrad(C.byref(pd), C.c_void_p(dw.data_ptr()), C.c_void_p(grad.data_ptr()), 1,
                                  _lib.stream_ptr()), "unpack")
    torch.cuda.synchronize()
    if transposed:
        inc = dw.view(k, k, cin, cout).permute(2, 3, 0, 1).flip(2, 3)
    else:
        inc = dw.view(k, k, cin, cout).permute(3, 2, 0, 1)
    assert torch.allclose(grad, before + inc, rtol=0, atol=1e-6)
    # slabs (deterministic split-K): partial gradients lying slab_elems apart are summed in index order
    parts = torch.randn((3,) + tuple(dw.shape), generator=g).to(dev)
    grad3 = torch.zeros_like(grad)
    _lib.check(L.gdn_unpack_wgrad_slabs(C.byref(pd), C.c_void_p(parts.data_ptr()), 3, C.c_int64(dw.numel()),
                                        C.c_void_p(grad3.data_ptr()), 0, _lib.stream_ptr()), "unpack slabs")
    torch.cuda.synchronize()
    tot = ((parts[0] + parts[1]) + parts[2])
    inc3 = tot.view(k, k, cin, cout).permute(2, 3, 0, 1).flip(2, 3) if transposed else tot.view(k, k, cin, cout).permute(3, 2, 0, 1)
    assert torch.equal(grad3, inc3.contiguous())


def test_pack_weights_table_matches_per_tensor_packs():
    """the one-launch job table (every packed tensor of a network) == the per-tensor gdn_pack_weights results"""
    from gdn_pytorch_b200.engine import PackDesc
    _lib, L = _L()
    g = torch.Generator().manual_seed(9)
    specs = [(64, 64, 9, False, False), (128, 64, 7, False, True), (512, 256, 3, False, False), (128, 256, 1, False, True),
             (256, 512, 4, True, False), (64, 128, 5, True, False), (256, 256, 5, False, True)]
    jsz = L.gdn_pack_job_size()
    blobs, cta0, outs, wants, keep = [], 0, [], [], []
    for cout, cin, k, transposed, scaled in specs:
        kk = k * k
        if transposed:
            w = torch.randn((cin, cout, k, k), generator=g).to(dev)
            pd = PackDesc(k, k, cout, cin, cout, cin, kk, cout * kk, k, 1, 1, 0)
        else:
            w = torch.randn((cout, cin, k, k), generator=g).to(dev)
            pd = PackDesc(k, k, cout, cin, cout, cin, cin * kk, kk, k, 1, 0, 0)
        sc = (torch.rand(cout, generator=g) + 0.5).to(dev) if scaled else None
        sp = C.c_void_p(sc.data_ptr() if sc is not None else None)
        want = torch.empty((kk, cout, cin), device=dev, dtype=torch.bfloat16)
        _lib.check(L.gdn_pack_weights(C.byref(pd), C.c_void_p(w.data_ptr()), sp, C.c_void_p(want.data_ptr()),
                                      _lib.stream_ptr()), "pack")
        out = torch.full((kk, cout, cin), float("nan"), device=dev, dtype=torch.bfloat16)
        buf, n = C.create_string_buffer(jsz), C.c_int(0)
        _lib.check(L.gdn_pack_job_fill(C.byref(pd), C.c_void_p(w.data_ptr()), sp, C.c_void_p(out.data_ptr()), cta0, buf,
                                       C.byref(n)), "job_fill")
        tb = 64 if kk <= 9 else (32 if kk <= 25 else 16)       # wide-tile version: 16 (a) x tb (b) tiles, pack_tile.cuh
        assert n.value in (((cin + 15) // 16) * ((cout + 15) // 16), ((cin + tb - 1) // tb) * ((cout + 15) // 16))
        blobs.append(buf.raw)
        cta0 += n.value
        outs.append(out)
        wants.append(want)
        keep += [w, sc]
    table = torch.frombuffer(bytearray(b"".join(blobs)), dtype=torch.uint8).to(dev)
    _lib.check(L.gdn_pack_weights_table(C.c_void_p(table.data_ptr()), len(blobs), cta0, 81, _lib.stream_ptr()), "table")
    torch.cuda.synchronize()
    for o, w_ in zip(outs, wants):
        assert torch.equal(o, w_)
    # an im2col'd thin layer is not tileable: the table builder reports it, the caller keeps gdn_pack_weights for it
    pd = PackDesc(9, 9, 64, 256, 64, 256, 3 * 81, 81, 9, 1, 0, 3)
    buf, n = C.create_string_buffer(jsz), C.c_int(0)
    w = torch.randn((64, 3, 9, 9), generator=g).to(dev)
    out = torch.empty((1, 64, 256), device=dev, dtype=torch.bfloat16)
    assert L.gdn_pack_job_fill(C.byref(pd), C.c_void_p(w.data_ptr()), C.c_void_p(None), C.c_void_p(out.data_ptr()), 0, buf,
                               C.byref(n)) != 0
