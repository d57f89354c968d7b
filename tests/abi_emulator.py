"""CPU emulation of a subset of the C ABI in include/gdn_b200.h (TEST INFRASTRUCTURE, never imported by the product).

Purpose: the engine (gdn_pytorch_b200/engine.py) is a *plan builder* -- it turns a layer graph into a list of
descriptor-driven library calls.  Whether that plan is right (buffers, offsets, weight layouts, accumulate flags,
reflection folds ...) does not depend on the GPU, so the CPU suite runs the plans against this emulator, which
implements the documented semantics of each entry point with torch CPU ops on the SAME descriptors (raw pointers
into CPU tensors).  bf16 operand rounding is reproduced (operands are read from bf16 buffers), accumulation is fp32.

Covered: gdn_bn_fold, gdn_bn_finalize, gdn_pack_weights, gdn_unpack_wgrad, gdn_im2col, gdn_head_gather, gdn_conv2d, gdn_conv2d_wgrad,
gdn_act_forward, gdn_bn_bwd_reduce, gdn_act_backward, gdn_act_backward_frozen, gdn_fold_grad, gdn_sqdiff_sum,
gdn_sqdiff_grad, gdn_tanh_chain_add.  Loss, metrics, Adam and the image kernels are not emulated (they have direct
GPU-vs-oracle tests).

Use:  with emulated_abi():  eng = Engine(..., device=torch.device("cpu")); run_ops(eng.fwd) ...
"""
import contextlib
import ctypes as C
import os

import numpy as np
import torch
import torch.nn.functional as F

_DT = {torch.float32: 4, torch.float64: 8, torch.bfloat16: 2, torch.float16: 2, torch.int32: 4}


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, C.c_void_p):
        return p.value or 0
    return int(p)


def _t(p, shape, dtype):
    """torch view (shared memory) of `shape` elements of `dtype` at raw address p"""
    a = _addr(p)
    assert a, "null pointer"
    n = int(np.prod(shape))
    buf = (C.c_char * (n * _DT[dtype])).from_address(a)
    return torch.frombuffer(buf, dtype=dtype, count=n).view(*shape)


def _obj(ref):
    return ref._obj if hasattr(ref, "_obj") else ref


class EmulatedLib:
    deterministic = 0      # set to 1 (before the engine is built) to emulate GDN_DETERMINISTIC=1: slab layout of the wgrads

    def __init__(self):
        self.calls = {}
        for name in dir(self):
            if name.startswith("_e_"):
                fn = getattr(self, name)
                setattr(self, "gdn_" + name[3:], self._counted("gdn_" + name[3:], fn))
        self._err = b""

    def _counted(self, name, fn):
        def call(*a):
            self.calls[name] = self.calls.get(name, 0) + 1
            return fn(*a)
        return call

    def __getattr__(self, name):
        if name.startswith("gdn_"):
            raise NotImplementedError("tests/abi_emulator.py does not emulate %s" % name)
        raise AttributeError(name)

    # ---- trivial entry points
    def _e_last_error(self):
        return self._err

    def _e_version(self):
        return 100

    def _e_sm_count(self):
        return 148

    def _e_deterministic(self):
        # the emulator always reduces in a fixed order; with deterministic = 1 it reproduces the LAYOUT of the deterministic
        # split-K (gdn_wgrad_desc.slabs / gdn_unpack_wgrad_slabs) so that the engine's plumbing is checked on the CPU
        return int(self.deterministic)

    # ---- BatchNorm fold (eval)
    def _e_bn_fold(self, gamma, beta, rmean, rvar, eps, scale, bias, c, s):
        eps = eps.value if hasattr(eps, "value") else eps
        g, b, m, v = (_t(p, (c,), torch.float32) for p in (gamma, beta, rmean, rvar))
        sc = g / torch.sqrt(v + eps)
        _t(scale, (c,), torch.float32).copy_(sc)
        _t(bias, (c,), torch.float32).copy_(b - m * sc)
        return 0

    # ---- weight pack
    def _e_pack_weights(self, pd, w, scale_a, out, s):
        k = _obj(pd)
        T = 1 if k.col_c else k.kh * k.kw
        res = torch.zeros((T, k.a_pad, k.b_pad), dtype=torch.float32)
        base = _addr(w)
        ai = torch.arange(k.a)

        def wread(idx):
            lo, hi = int(idx.min()), int(idx.max())
            flat = _t(base + 4 * lo, (hi - lo + 1,), torch.float32)
            return flat[(idx - lo).reshape(-1)].reshape(idx.shape)
        if k.col_c:
            for tt in range(k.kh * k.kw):
                r, s2 = divmod(tt, k.kw)
                if k.flip:
                    r, s2 = k.kh - 1 - r, k.kw - 1 - s2
                for c in range(k.col_c):
                    res[0, :k.a, tt * k.col_c + c] = wread(ai * k.stride_a + c * k.stride_b + r * k.stride_r + s2 * k.stride_s)
        else:
            bi = torch.arange(k.b)
            for t in range(T):
                r, s2 = divmod(t, k.kw)
                if k.flip:
                    r, s2 = k.kh - 1 - r, k.kw - 1 - s2
                idx = ai[:, None] * k.stride_a + bi[None, :] * k.stride_b + r * k.stride_r + s2 * k.stride_s
                res[t, :k.a, :k.b] = wread(idx)
        if _addr(scale_a):
            res[:, :k.a, :] *= _t(scale_a, (k.a,), torch.float32)[None, :, None]
        _t(out, (T, k.a_pad, k.b_pad), torch.bfloat16).copy_(res.to(torch.bfloat16))
        return 0

    # ---- thin-layer im2col
    def _e_im2col(self, src, dst, n, c, h, w, kh, kw, pad, reflect, kpad, s):
        x = _t(src, (n, c, h, w), torch.float32)
        xp = F.pad(x, (pad,) * 4, mode="reflect" if reflect else "constant")
        cols = torch.zeros((n, h, w, kpad), dtype=torch.float32)
        for r in range(kh):
            for q in range(kw):
                for ch in range(c):
                    cols[..., (r * kw + q) * c + ch] = xp[:, ch, r:r + h, q:q + w]
        _t(dst, (n, h, w, kpad), torch.bfloat16).copy_(cols.to(torch.bfloat16))
        return 0

    # ---- 64 -> 1 heads: shifted sum of the per-tap 1x1 results
    def _e_head_gather(self, z, z_is_half, zc, n, h, w, k, pad, tanh_out, out, s):
        zt = _t(z, (n, h, w, zc), torch.float16 if z_is_half else torch.bfloat16).float()
        zp = F.pad(zt.permute(0, 3, 1, 2), (pad, k - 1 - pad, pad, k - 1 - pad))      # zero rows/cols = skipped taps
        acc = torch.zeros((n, h, w), dtype=torch.float32)
        for r in range(k):
            for q in range(k):
                acc += zp[:, r * k + q, r:r + h, q:q + w]
        _t(out, (n, h, w), torch.float32).copy_(torch.tanh(acc) if tanh_out else acc)
        return 0

    # ---- implicit-GEMM convolution (semantics of the gdn_conv_desc comment)
    def _e_conv2d(self, dref, s):
        d = _obj(dref)

        def phys(a):
            return _t(a.ptr, (a.n, a.h + 2 * a.pad, a.w + 2 * a.pad, a.c), torch.bfloat16).float()
        X, p = phys(d.src0), d.src0.pad
        if d.src1.ptr:
            assert (d.src1.h, d.src1.w, d.src1.pad) == (d.src0.h, d.src0.w, d.src0.pad)
            X = torch.cat((X, phys(d.src1)), -1)
        n, Hp, Wp, cin = X.shape
        if cin % 64:
            return -2
        kh, kw, st = d.kh, d.kw, d.stride
        Wt = _t(d.weights, (kh * kw, d.cout_pad, cin), torch.bfloat16).float()
        Wc = Wt.permute(1, 2, 0).reshape(d.cout_pad, cin, kh, kw)
        ylo, xlo = d.off_y + p, d.off_x + p
        yhi, xhi = (d.out_h - 1) * st + kh - 1 + ylo, (d.out_w - 1) * st + kw - 1 + xlo
        pt, pl = max(0, -ylo), max(0, -xlo)
        pb, pr = max(0, yhi - (Hp - 1)), max(0, xhi - (Wp - 1))
        Xn = F.pad(X.permute(0, 3, 1, 2), (pl, pr, pt, pb))
        crop = Xn[:, :, ylo + pt: yhi + pt + 1, xlo + pl: xhi + pl + 1]
        acc = F.conv2d(crop, Wc, stride=st)[:, :d.cout].permute(0, 2, 3, 1).contiguous()   # (n, out_h, out_w, cout)
        assert acc.shape[1:3] == (d.out_h, d.out_w)
        if d.stat_sum and not d.bwd_raw:
            _t(d.stat_sum, (d.cout,), torch.float64).add_(acc.double().sum((0, 1, 2)))
            _t(d.stat_sqsum, (d.cout,), torch.float64).add_((acc.double() ** 2).sum((0, 1, 2)))
            if d.fin_counter:      # fused BatchNorm finalisation: what the last CTA does with the finished sums
                assert int(_t(d.fin_counter, (1,), torch.int32)[0]) == 0, "fin_counter must be 0 before the launch"
                self._e_bn_finalize(d.stat_sum, d.stat_sqsum, d.fin_count, d.fin_gamma, d.fin_beta, d.fin_eps, d.fin_momentum,
                                    d.fin_running_mean, d.fin_running_var, d.fin_scale, d.fin_shift, d.fin_mean, d.fin_rstd,
                                    d.fin_coef4, d.cout, s)
        v = acc
        if d.bias:
            v = v + _t(d.bias, (d.cout,), torch.float32)
        if d.relu:
            v = torch.relu(v)
        sy, sx = d.dst_sy, d.dst_sx
        ys = slice(d.dst_oy, d.dst_oy + (d.out_h - 1) * sy + 1, sy)
        xs = slice(d.dst_ox, d.dst_ox + (d.out_w - 1) * sx + 1, sx)
        assert d.dst_oy + (d.out_h - 1) * sy < d.dst_h and d.dst_ox + (d.out_w - 1) * sx < d.dst_w
        if d.resid:
            v = v + _t(d.resid, (n, d.dst_h, d.dst_w, d.cout), torch.float32)[:, ys, xs]
        if d.bwd_raw:
            # fused BatchNorm(+ReLU)-backward statistics of the unit that produced the destination tensor (header comment)
            assert d.cout == d.cout_pad and d.cout % 32 == 0 and not d.tanh_out and not d.out16_is_half
            raw = _t(d.bwd_raw, (n, d.dst_h, d.dst_w, d.cout), torch.float16)[:, ys, xs].float()
            cf = _t(d.bwd_coef, (d.cout, 4), torch.float32)
            if d.bwd_relu:
                v = v * ((raw * cf[:, 0] + cf[:, 1]) > 0)
            xhat = (raw - cf[:, 2]) * cf[:, 3]
            _t(d.stat_sum, (d.cout,), torch.float64).add_(v.double().sum((0, 1, 2)))
            _t(d.stat_sqsum, (d.cout,), torch.float64).add_((v * xhat).double().sum((0, 1, 2)))
        if d.tanh_out:
            v = torch.tanh(v)
        if d.out_f32:
            _t(d.out_f32, (n, d.dst_h, d.dst_w, d.cout), torch.float32)[:, ys, xs] = v
        if d.out_bf16.ptr:
            ob = d.out_bf16
            assert (ob.h, ob.w, ob.c) == (d.dst_h, d.dst_w, d.cout)
            dt = torch.float16 if d.out16_is_half else torch.bfloat16
            buf = _t(ob.ptr, (n, ob.h + 2 * ob.pad, ob.w + 2 * ob.pad, ob.c), dt)
            q = ob.pad
            buf[:, q + d.dst_oy: q + d.dst_oy + (d.out_h - 1) * sy + 1: sy,
                q + d.dst_ox: q + d.dst_ox + (d.out_w - 1) * sx + 1: sx] = v.to(dt)
            if d.out_reflect and q:
                inner = buf[:, q:q + ob.h, q:q + ob.w].float().permute(0, 3, 1, 2)
                buf.copy_(F.pad(inner, (q,) * 4, mode="reflect").permute(0, 2, 3, 1).to(dt))
        return 0

    # ---- BatchNorm apply / ReLU / residual / halo / upsample / dilation
    def _e_act_forward(self, aref, s):
        a = _obj(aref)
        shp = (a.n, a.h, a.w, a.c)
        if a.src_f32:
            y = _t(a.src_f32, shp, torch.float32).clone()
        else:
            y = _t(a.src_bf16, shp, torch.float16 if a.src16_is_half else torch.bfloat16).float()
        if a.scale:
            y = y * _t(a.scale, (a.c,), torch.float32) + _t(a.shift, (a.c,), torch.float32)
        if a.relu:
            y = torch.relu(y)
        if a.resid:
            y = y + _t(a.resid, shp, torch.float32)
        if a.out_f32:
            _t(a.out_f32, shp, torch.float32).copy_(y)
        if a.out_bf16:
            z = y.permute(0, 3, 1, 2)
            if a.up:
                z = F.interpolate(z, scale_factor=2, mode="bilinear", align_corners=(a.up == 2))
            elif a.dilate:
                zz = torch.zeros((a.n, a.c, 2 * a.h, 2 * a.w))
                zz[:, :, ::2, ::2] = z
                z = zz
            OH, OW, q = z.shape[2], z.shape[3], a.pad
            buf = _t(a.out_bf16, (a.n, OH + 2 * q, OW + 2 * q, a.c), torch.bfloat16)
            if a.reflect and q:
                buf.copy_(F.pad(z, (q,) * 4, mode="reflect").permute(0, 2, 3, 1).to(torch.bfloat16))
            else:
                buf[:, q:q + OH, q:q + OW] = z.permute(0, 2, 3, 1).to(torch.bfloat16)
        return 0

    def _e_act_backward_frozen(self, bref, s):
        b = _obj(bref)
        shp = (b.n, b.h, b.w, b.c)
        g = _t(b.dact, shp, torch.float32).clone()
        if b.relu:
            y = _t(b.y_f32, shp, torch.float32) if b.y_f32 else _t(b.y_bf16, shp, torch.bfloat16).float()
            g = g * (y > 0)
        if b.scale:
            g = g * _t(b.scale, (b.c,), torch.float32)
        _t(b.dy, shp, torch.bfloat16).copy_(g.to(torch.bfloat16))
        return 0

    def _e_fold_grad(self, fref, s):
        f = _obj(fref)
        sc = 2 if (f.up or f.dilate) else 1
        Hq, Wq = f.h * sc + 2 * f.pad, f.w * sc + 2 * f.pad
        dpad = _t(f.dpad, (f.n, Hq, Wq, f.ctot), torch.bfloat16 if f.dpad_is_bf16 else torch.float32).float()
        dpad = dpad[..., f.c_off:f.c_off + f.c].permute(0, 3, 1, 2)
        with torch.enable_grad():             # the adjoint of the input transform, by autograd of the transform itself
            x = torch.zeros((f.n, f.c, f.h, f.w), requires_grad=True)
            z = x
            if f.up:
                z = F.interpolate(z, scale_factor=2, mode="bilinear", align_corners=(f.up == 2))
            elif f.dilate:
                zz = torch.zeros((f.n, f.c, 2 * f.h, 2 * f.w))
                zz[:, :, ::2, ::2] = z
                z = zz
            if f.pad:
                z = F.pad(z, (f.pad,) * 4, mode="reflect" if f.reflect else "constant")
            (gx,) = torch.autograd.grad(z, x, dpad.contiguous())
        out = _t(f.dact, (f.n, f.h, f.w, f.c), torch.float32)
        gx = gx.permute(0, 2, 3, 1)
        if f.accumulate:
            out.add_(gx)
        else:
            out.copy_(gx)
        return 0

    # ---- training mode: batch statistics, BatchNorm backward, weight gradients
    def _e_bn_finalize(self, sum_, sqsum, count, gamma, beta, eps, momentum, rmean, rvar, scale, shift, mean, rstd, coef4,
                       c, s):
        count, eps, momentum = (v.value if hasattr(v, "value") else v for v in (count, eps, momentum))
        m = _t(sum_, (c,), torch.float64) / count
        var = _t(sqsum, (c,), torch.float64) / count - m * m
        rs = 1.0 / torch.sqrt(var + eps)
        g, b = _t(gamma, (c,), torch.float32).double(), _t(beta, (c,), torch.float32).double()
        _t(scale, (c,), torch.float32).copy_((g * rs).float())
        _t(shift, (c,), torch.float32).copy_((b - m * g * rs).float())
        _t(mean, (c,), torch.float32).copy_(m.float())
        _t(rstd, (c,), torch.float32).copy_(rs.float())
        if _addr(coef4):
            _t(coef4, (c, 4), torch.float32).copy_(torch.stack([(g * rs).float(), (b - m * g * rs).float(), m.float(),
                                                                  rs.float()], 1))
        if _addr(rmean):
            rm, rv = _t(rmean, (c,), torch.float32), _t(rvar, (c,), torch.float32)
            rm.mul_(1 - momentum).add_((momentum * m).float())
            rv.mul_(1 - momentum).add_((momentum * var * count / (count - 1)).float())
        return 0

    def _bn_bwd_terms(self, b):
        shp = (b.n, b.h, b.w, b.c)
        g = _t(b.dact, shp, torch.bfloat16).float() if b.dact_is_bf16 else _t(b.dact, shp, torch.float32).clone()
        raw = _t(b.raw, shp, torch.float16 if b.raw_is_half else torch.bfloat16).float()
        sc, sh = _t(b.scale, (b.c,), torch.float32), _t(b.shift, (b.c,), torch.float32)
        if b.relu:
            g = g * ((raw * sc + sh) > 0)
        xhat = (raw - _t(b.mean, (b.c,), torch.float32)) * _t(b.rstd, (b.c,), torch.float32)
        return g, xhat, sc

    def _e_bn_bwd_reduce(self, bref, s):
        b = _obj(bref)
        g, xhat, _ = self._bn_bwd_terms(b)
        _t(b.sum_g, (b.c,), torch.float64).add_(g.double().sum((0, 1, 2)))
        _t(b.sum_gx, (b.c,), torch.float64).add_((g * xhat).double().sum((0, 1, 2)))
        return 0

    def _e_act_backward(self, bref, s):
        b = _obj(bref)
        g, xhat, sc = self._bn_bwd_terms(b)
        npix = float(b.n * b.h * b.w)
        sg, sgx = _t(b.sum_g, (b.c,), torch.float64), _t(b.sum_gx, (b.c,), torch.float64)
        dy = sc * (g - (sg / npix).float() - xhat * (sgx / npix).float())
        assert not b.dilate
        _t(b.dy, (b.n, b.h, b.w, b.c), torch.bfloat16).copy_(dy.to(torch.bfloat16))
        if b.dgamma:
            _t(b.dgamma, (b.c,), torch.float32).add_(sgx.float())
            _t(b.dbeta, (b.c,), torch.float32).add_(sg.float())
        return 0

    def _e_conv2d_wgrad(self, wref, s):
        w = _obj(wref)

        def phys(a):
            return _t(a.ptr, (a.n, a.h + 2 * a.pad, a.w + 2 * a.pad, a.c), torch.bfloat16).float()
        X, p = phys(w.x0), w.x0.pad
        if w.x1.ptr:
            X = torch.cat((X, phys(w.x1)), -1)
        n, Hp, Wp, cin = X.shape
        dy = _t(w.dy.ptr, (n, w.out_h, w.out_w, w.cout_pad), torch.bfloat16).float()
        assert (w.dy.h, w.dy.w, w.dy.c, w.dy.pad) == (w.out_h, w.out_w, w.cout_pad, 0)
        st = w.stride
        ylo, xlo = w.off_y + p, w.off_x + p
        yhi, xhi = (w.out_h - 1) * st + w.kh - 1 + ylo, (w.out_w - 1) * st + w.kw - 1 + xlo
        pt, pl = max(0, -ylo), max(0, -xlo)
        pb, pr = max(0, yhi - (Hp - 1)), max(0, xhi - (Wp - 1))
        Xn = F.pad(X.permute(0, 3, 1, 2), (pl, pr, pt, pb)).permute(0, 2, 3, 1)
        if w.slabs:
            # deterministic split-K: split s STORES its partial to slab s (no accumulation, dw untouched).  Emulated with
            # min(max_slabs, 3, images) splits over the images; the count goes to *splits_used (host memory)
            assert w.max_slabs >= 1 and bool(w.splits_used), "slabs need max_slabs and splits_used"
            nsp = max(1, min(w.max_slabs, 3, n))
            w.splits_used[0] = nsp
            elems = w.kh * w.kw * cin * w.cout_pad
            slabs = _t(w.slabs, (nsp, w.kh * w.kw, cin, w.cout_pad), torch.float32)
            slabs.fill_(float("nan"))             # every element of every used slab must be written
            bounds = [n * i // nsp for i in range(nsp + 1)]
            for sp in range(nsp):
                lo, hi = bounds[sp], bounds[sp + 1]
                for r in range(w.kh):
                    for q in range(w.kw):
                        y0, x0 = ylo + pt + r, xlo + pl + q
                        xs = Xn[lo:hi, y0: y0 + (w.out_h - 1) * st + 1: st, x0: x0 + (w.out_w - 1) * st + 1: st]
                        slabs[sp, r * w.kw + q] = torch.einsum("nyxi,nyxo->io", xs, dy[lo:hi])
            assert elems == slabs[0].numel()
            return 0
        dw = _t(w.dw, (w.kh * w.kw, cin, w.cout_pad), torch.float32)
        for r in range(w.kh):
            for q in range(w.kw):
                y0, x0 = ylo + pt + r, xlo + pl + q
                xs = Xn[:, y0: y0 + (w.out_h - 1) * st + 1: st, x0: x0 + (w.out_w - 1) * st + 1: st]
                dw[r * w.kw + q] += torch.einsum("nyxi,nyxo->io", xs, dy)
        return 0

    def _e_unpack_wgrad_slabs(self, pd, dw, slabs, slab_elems, grad, accumulate, s):
        """sum `slabs` partial gradients lying slab_elems floats apart in index order, then scatter like gdn_unpack_wgrad"""
        k = _obj(pd)
        T = 1 if k.col_c else k.kh * k.kw
        slab_elems = slab_elems.value if hasattr(slab_elems, "value") else slab_elems
        assert slabs >= 1 and (slabs == 1 or slab_elems >= T * k.b_pad * k.a_pad)
        tot = _t(dw, (T, k.b_pad, k.a_pad), torch.float32).clone()
        for sl in range(1, slabs):
            tot += _t(_addr(dw) + 4 * sl * slab_elems, (T, k.b_pad, k.a_pad), torch.float32)
        assert torch.isfinite(tot).all(), "a slab element was never written"
        return self._e_unpack_wgrad(pd, tot.data_ptr(), grad, accumulate, s, _keep=tot)

    def _e_unpack_wgrad(self, pd, dw, grad, accumulate, s, _keep=None):
        k = _obj(pd)
        T = 1 if k.col_c else k.kh * k.kw
        src = _t(dw, (T, k.b_pad, k.a_pad), torch.float32)
        base = _addr(grad)
        ai = torch.arange(k.a)

        def scatter(idx, vals):
            lo, hi = int(idx.min()), int(idx.max())
            flat = _t(base + 4 * lo, (hi - lo + 1,), torch.float32)
            ii = (idx - lo).reshape(-1)
            if accumulate:
                flat.index_add_(0, ii, vals.reshape(-1))
            else:
                flat[ii] = vals.reshape(-1)
        if k.col_c:
            for tt in range(k.kh * k.kw):
                r, s2 = divmod(tt, k.kw)
                if k.flip:
                    r, s2 = k.kh - 1 - r, k.kw - 1 - s2
                for c in range(k.col_c):
                    scatter(ai * k.stride_a + c * k.stride_b + r * k.stride_r + s2 * k.stride_s,
                            src[0, tt * k.col_c + c, :k.a])
        else:
            bi = torch.arange(k.b)
            for t in range(T):
                r, s2 = divmod(t, k.kw)
                if k.flip:
                    r, s2 = k.kh - 1 - r, k.kw - 1 - s2
                idx = ai[:, None] * k.stride_a + bi[None, :] * k.stride_b + r * k.stride_r + s2 * k.stride_s
                scatter(idx, src[t, :k.b, :k.a].t())
        return 0

    # ---- guidance-loss helpers
    def _e_sqdiff_sum(self, a, b, n, out, s):
        n = n.value if hasattr(n, "value") else n
        d = _t(a, (n,), torch.float32).double() - _t(b, (n,), torch.float32).double()
        _t(out, (1,), torch.float64).add_((d * d).sum())
        return 0

    def _e_sqdiff_grad(self, a, b, n, coef, grad, s):
        n = n.value if hasattr(n, "value") else n
        coef = coef.value if hasattr(coef, "value") else coef
        _t(grad, (n,), torch.float32).copy_(coef * (_t(a, (n,), torch.float32) - _t(b, (n,), torch.float32)))
        return 0

    def _e_tanh_chain_add(self, dout, out, n, scale, dpre, s):
        n = n.value if hasattr(n, "value") else n
        scale = scale.value if hasattr(scale, "value") else scale
        o = _t(out, (n,), torch.float32)
        _t(dpre, (n,), torch.float32).add_(scale * _t(dout, (n,), torch.float32) * (1 - o * o))
        return 0


@contextlib.contextmanager
def emulated_abi():
    """route gdn_pytorch_b200._lib.lib() to the emulator; autotuning and the batched pack tables are switched off
    (they are GPU-side optimisations of the same calls)"""
    from gdn_pytorch_b200 import _lib
    saved = _lib._lib
    env = {k: os.environ.get(k) for k in ("GDN_AUTOTUNE", "GDN_PACK_TABLE", "GDN_SIDE")}
    os.environ["GDN_AUTOTUNE"] = "0"
    os.environ["GDN_PACK_TABLE"] = "0"
    os.environ["GDN_SIDE"] = "0"              # no CUDA side streams: the plan runs in list order
    emu = EmulatedLib()
    _lib._lib = emu
    try:
        yield emu
    finally:
        _lib._lib = saved
        for k, v in env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def run_ops(ops):
    for op in ops:
        op(None)


def engine_forward(eng, x):
    """Engine.forward without the CUDA-only argument checks: x is a CPU fp32 NCHW tensor"""
    eng._x = x.contiguous()
    run_ops(eng.pack_ops)
    if eng.do_bwd:
        run_ops(eng.pack_ops_bwd)
    run_ops(eng.fwd)
