"""Python surface of the fused loss / metric / optimizer kernels.

  compute_errors(gt_np, gt, pred, crop=True)   drop-in for calculate_error.compute_errors
                                               (/root/reference/src/calculate_error.py:10-103): list of 8 floats
  compute_errors_NYU / compute_errors_Make3D   the other two evaluation protocols (:105-182), same kernel
  eigen_metrics_device(...)                    same numbers without the host sync (+ exact integer delta counts)
  LossKernels                                  RtoD / DtoD training loss with analytic gradient
                                               (/root/reference/src/trainer.py:433-456, 705-757)
  FusedAdam                                    optim.Adam(params, lr, betas, eps, weight_decay) replacement
                                               (/root/reference/src/GDN_main.py:157,173,184)
All of them call libgdn_b200.so; there is no PyTorch fallback.
"""
import ctypes as C

import torch

from . import _lib


class LossDesc(C.Structure):
    _fields_ = [("out", C.c_void_p), ("gt", C.c_void_p), ("sparse", C.c_void_p), ("sparse_stride", C.c_int64),
                ("rgb", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("maxabs", C.c_void_p),
                ("mode", C.c_int32), ("sums", C.c_void_p), ("dout", C.c_void_p), ("dpre", C.c_void_p),
                ("grad_scale", C.c_float)]


def _chk_map(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32):
        raise RuntimeError("gdn_b200: %s must be a CUDA fp32 tensor (no CPU fallback)" % name)
    return t.contiguous()


# ------------------------------------------------------------------------------------------------ metrics
def eigen_metrics_device(gt_np, gt, pred, crop=True):
    """-> (out8: float64[8] cuda, counts: int64[B,4] cuda); no host synchronisation."""
    gt_np, gt, pred = _chk_map(gt_np, "gt_np"), _chk_map(gt, "gt"), _chk_map(pred, "pred")
    B, H, W = gt.shape[0], gt.shape[-2], gt.shape[-1]
    if gt.numel() != B * H * W or pred.numel() != B * H * W:
        raise ValueError("gdn_b200.compute_errors expects single-channel (B,1,H,W) depth maps")
    if gt_np.numel() != B * H * W:
        gt_np = gt_np[:, 0].contiguous()   # the reference only reads channel 0 (calculate_error.py:35)
    out8 = torch.zeros(8, dtype=torch.float64, device=gt.device)
    counts = torch.zeros((B, 4), dtype=torch.int64, device=gt.device)
    L = _lib.lib()
    ws = torch.empty(L.gdn_depth_metrics_workspace_bytes(B) // 8, dtype=torch.int64, device=gt.device)
    with torch.cuda.device(gt.device):
        rc = L.gdn_eigen_metrics(C.c_void_p(gt_np.data_ptr()), C.c_void_p(gt.data_ptr()), C.c_void_p(pred.data_ptr()),
                                 B, H, W, int(bool(crop)), C.c_void_p(out8.data_ptr()), C.c_void_p(counts.data_ptr()),
                                 C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel() * 8), _lib.stream_ptr())
    _lib.check(rc, "eigen_metrics")
    return out8, counts


def compute_errors(gt_np, gt, pred, crop=True):
    """[abs_diff, abs_rel, sq_rel, a1, a2, a3, rmse, rmse_log] as Python floats (one host read instead of the
    reference's per-image syncs)."""
    out8, _ = eigen_metrics_device(gt_np, gt, pred, crop)
    return out8.tolist()


def _depth_metrics(variant, gt_np, gt, pred, crop):
    gt, pred = _chk_map(gt, "gt"), _chk_map(pred, "pred")
    B, H, W = gt.shape[0], gt.shape[-2], gt.shape[-1]
    if gt.numel() != B * H * W or pred.numel() != B * H * W:
        raise ValueError("gdn_b200 depth metrics expect single-channel (B,1,H,W) depth maps")
    if gt_np is not None:
        gt_np = _chk_map(gt_np, "gt_np")
        if gt_np.numel() != B * H * W:
            gt_np = gt_np[:, 0].contiguous()
    out8 = torch.zeros(8, dtype=torch.float64, device=gt.device)
    counts = torch.zeros((B, 4), dtype=torch.int64, device=gt.device)
    L = _lib.lib()
    ws = torch.empty(L.gdn_depth_metrics_workspace_bytes(B) // 8, dtype=torch.int64, device=gt.device)
    with torch.cuda.device(gt.device):
        rc = L.gdn_depth_metrics(variant, C.c_void_p(gt_np.data_ptr() if gt_np is not None else None),
                                 C.c_void_p(gt.data_ptr()), C.c_void_p(pred.data_ptr()), B, H, W, int(bool(crop)),
                                 C.c_void_p(out8.data_ptr()), C.c_void_p(counts.data_ptr()),
                                 C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel() * 8), _lib.stream_ptr())
    _lib.check(rc, "depth_metrics")
    return out8, counts


def compute_errors_NYU(gt, pred, crop=True):
    """drop-in for calculate_error.compute_errors_NYU (/root/reference/src/calculate_error.py:105-151):
    [abs_diff, abs_rel, log10, a1, a2, a3, rmse, rmse_log] as Python floats"""
    return _depth_metrics(1, None, gt, pred, crop)[0].tolist()


def compute_errors_Make3D(gt_np, gt, pred):
    """drop-in for calculate_error.compute_errors_Make3D (:153-182): [abs_diff, abs_rel, ave_log10, rmse]"""
    o = _depth_metrics(2, gt_np, gt, pred, False)[0].tolist()
    return [o[0], o[1], o[2], o[6]]


# --------------------------------------------------------------------------------------------------- loss
class LossKernels:
    """Workspace + launch wrappers for the fused training loss.  Results stay on the device:
    ``terms`` = float64[8]: [berhu_sum, second_sum, sqdiff_sum, latent1..4 sums, unused]."""
    LATENT_W = (1.0, 2.5, 14.0, 12.0)

    def __init__(self, device):
        self.device = device
        self.maxabs = torch.zeros(1, dtype=torch.float32, device=device)
        self.terms = torch.zeros(8, dtype=torch.float64, device=device)
        self.L = _lib.lib()

    def absdiff_max(self, out, gt):
        self.maxabs.zero_()
        rc = self.L.gdn_absdiff_max(C.c_void_p(out.data_ptr()), C.c_void_p(gt.data_ptr()), C.c_int64(out.numel()),
                                    C.c_void_p(self.maxabs.data_ptr()), _lib.stream_ptr())
        _lib.check(rc, "absdiff_max")
        return self.maxabs

    def loss(self, mode, out, gt, sparse, rgb, dout=None, dpre=None, grad_scale=1.0):
        """mode 0 = RtoD (BerHu + edge-aware smoothness), 1 = DtoD (BerHu + 3*Sobel).  absdiff_max() (and, when the
        batch is sharded, an all-reduce MAX of self.maxabs) must have run before."""
        N, H, W = out.shape[0], out.shape[-2], out.shape[-1]
        # the kernel indexes every operand with the prediction's extents: a mismatched tensor would be read out of bounds
        for name, t, chans in (("gt", gt, (1,)), ("sparse", sparse, (1, 3)), ("rgb", rgb, (3,)), ("dout", dout, (1,)),
                               ("dpre", dpre, (1,))):
            if t is None:
                continue
            if t.shape[0] != N or t.shape[-2:] != out.shape[-2:] or t.numel() // (N * H * W) not in chans or not t.is_contiguous():
                raise ValueError("gdn_b200 loss: %s has shape %s, expected (%d, %s, %d, %d) contiguous"
                                 % (name, tuple(t.shape), N, "|".join(map(str, chans)), H, W))
        d = LossDesc()
        d.out, d.gt = out.data_ptr(), gt.data_ptr()
        if sparse is not None:
            d.sparse = sparse.data_ptr()
            d.sparse_stride = sparse.stride(0)
        d.rgb = rgb.data_ptr() if rgb is not None else None
        d.n, d.h, d.w = N, H, W
        d.maxabs = self.maxabs.data_ptr()
        d.mode = mode
        d.sums = self.terms.data_ptr()
        d.dout = dout.data_ptr() if dout is not None else None
        d.dpre = dpre.data_ptr() if dpre is not None else None
        d.grad_scale = grad_scale
        self.terms.zero_()
        _lib.check(self.L.gdn_loss(C.byref(d), _lib.stream_ptr()), "loss")

    def latent(self, feats, feats_tar):
        """accumulate the four feature squared-difference sums into terms[3:7]"""
        for i, (a, b) in enumerate(zip(feats, feats_tar)):
            rc = self.L.gdn_sqdiff_sum(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_int64(a.numel()),
                                       C.c_void_p(self.terms.data_ptr() + 8 * (3 + i)), _lib.stream_ptr())
            _lib.check(rc, "sqdiff_sum")

    def latent_grad(self, feats, feats_tar, grads):
        """opt-in guidance gradient: grads[i] <- d latent_loss / d feats[i]
        = 1.5/4 * w_i * 2/numel_i * (feats[i] - feats_tar[i])  (latent_loss of trainer.py:726-733 WITHOUT the
        no_grad of :699-703)"""
        for w, a, b, g in zip(self.LATENT_W, feats, feats_tar, grads):
            coef = 1.5 / 4.0 * w * 2.0 / float(a.numel())
            rc = self.L.gdn_sqdiff_grad(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_int64(a.numel()),
                                        C.c_float(coef), C.c_void_p(g.data_ptr()), _lib.stream_ptr())
            _lib.check(rc, "sqdiff_grad")

    def tanh_chain_add(self, dout, out, dpre, scale=1.0):
        """dpre += scale * dout * (1 - out^2): add a gradient w.r.t. the network output to dL/d(pre-tanh)"""
        rc = self.L.gdn_tanh_chain_add(C.c_void_p(dout.data_ptr()), C.c_void_p(out.data_ptr()), C.c_int64(out.numel()),
                                       C.c_float(scale), C.c_void_p(dpre.data_ptr()), _lib.stream_ptr())
        _lib.check(rc, "tanh_chain_add")

    def assemble(self, mode, npix, feat_numels=None):
        """device-side scalar assembly (tiny fp64 tensor ops, no sync) -> dict of 0-dim tensors"""
        t = self.terms
        out_loss = 3.0 * t[0] / npix
        res = {"output_loss": out_loss, "rmse_loss": torch.sqrt(t[2] / npix), "c": 0.2 * self.maxabs[0]}
        if mode == 0:
            smooth = 0.1 * t[1] / npix
            lat = torch.zeros((), dtype=torch.float64, device=t.device)
            if feat_numels is not None:
                for i, (w, n) in enumerate(zip(self.LATENT_W, feat_numels)):
                    lat = lat + w * t[3 + i] / n
                lat = 1.5 * (lat / 4)
            res.update({"smooth_loss": smooth, "latent_loss": lat, "loss": out_loss + lat + smooth})
        else:
            grad_loss = 3.0 * t[1] / npix
            res.update({"gradient_loss": grad_loss, "loss": out_loss + grad_loss})
        return res


# --------------------------------------------------------------------------------------------------- Adam
class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (coupled L2 weight decay, bias correction) in one kernel per tensor -- or ONE
    kernel for the whole model when parameters and gradients live in flat buffers (``flat=(params, grads)``,
    set up by trainer.flatten_parameters)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, flat=None):
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.flat = flat
        self.dyn = None            # optional device float[2] = (lr, steps taken), see enable_device_step()
        self._flat_state = None
        self._step = 0
        self.grad_scale = 1.0
        self.L = _lib.lib()

    def enable_device_step(self):
        """keep (lr, step count) in a 2-float device buffer: nothing step-dependent is passed by value any more,
        so the optimizer launch can be captured in a CUDA graph"""
        if self.dyn is None:
            self.dyn = torch.tensor([self.param_groups[0]["lr"], float(self._step)], dtype=torch.float32,
                                    device=self.flat[0].device)
        return self.dyn

    def set_lr(self, lr):
        for g in self.param_groups:
            g["lr"] = lr
        if self.dyn is not None:
            self.dyn[0:1].fill_(lr)

    def set_hyper(self, betas, eps, weight_decay):
        """betas / eps / weight decay are passed by value at launch: valid until the step is captured in a CUDA graph"""
        for g in self.param_groups:
            g["betas"], g["eps"], g["weight_decay"] = tuple(betas), eps, weight_decay

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self._step += 1
        L = self.L
        if self.flat is not None:
            fp, fg = self.flat
            if self._flat_state is None:
                self._flat_state = (torch.zeros_like(fp), torch.zeros_like(fp))
            m, v = self._flat_state
            g = self.param_groups[0]
            if self.dyn is not None:
                # step-dependent scalars live in device memory (written by set_dyn() outside a captured graph)
                rc = L.gdn_adam_step_dyn(C.c_void_p(fp.data_ptr()), C.c_void_p(fg.data_ptr()), C.c_void_p(m.data_ptr()),
                                         C.c_void_p(v.data_ptr()), C.c_int64(fp.numel()), C.c_void_p(self.dyn.data_ptr()),
                                         C.c_double(g["betas"][0]), C.c_double(g["betas"][1]), C.c_float(g["eps"]),
                                         C.c_float(g["weight_decay"]), C.c_float(self.grad_scale), _lib.stream_ptr())
                _lib.check(rc, "adam_step_dyn(flat)")
                return loss
            rc = L.gdn_adam_step(C.c_void_p(fp.data_ptr()), C.c_void_p(fg.data_ptr()), C.c_void_p(m.data_ptr()),
                                 C.c_void_p(v.data_ptr()), C.c_int64(fp.numel()), C.c_float(g["lr"]),
                                 C.c_float(g["betas"][0]), C.c_float(g["betas"][1]), C.c_float(g["eps"]),
                                 C.c_float(g["weight_decay"]), self._step, C.c_float(self.grad_scale), _lib.stream_ptr())
            _lib.check(rc, "adam_step(flat)")
            return loss
        for grp in self.param_groups:
            for p in grp["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("gdn_b200.FusedAdam: CUDA fp32 parameters only")
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["step"] = 0
                st["step"] += 1
                g = p.grad.contiguous()
                rc = L.gdn_adam_step(C.c_void_p(p.data_ptr()), C.c_void_p(g.data_ptr()),
                                     C.c_void_p(st["exp_avg"].data_ptr()), C.c_void_p(st["exp_avg_sq"].data_ptr()),
                                     C.c_int64(p.numel()), C.c_float(grp["lr"]), C.c_float(grp["betas"][0]),
                                     C.c_float(grp["betas"][1]), C.c_float(grp["eps"]), C.c_float(grp["weight_decay"]),
                                     st["step"], C.c_float(self.grad_scale), _lib.stream_ptr())
                _lib.check(rc, "adam_step")
                p.add_(0)  # bump the version counter: engines re-pack their bf16 weight copies on the next forward
        return loss
