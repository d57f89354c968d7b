"""Execution engine: turns a layer graph (graph.py) into a static list of sm_100a kernel launches.

Data layout in HBM (per engine instance, all buffers allocated once):
  * activations: NHWC bf16, one buffer per (tensor, variant) where the variant is what the consuming conv needs as
    its TMA source: plain / reflection border of P pixels / x2 bilinear upsample (+ border) / zero-dilated x2.
    Zero padding is never materialised -- TMA out-of-bounds fill supplies it.
  * residual stream: fp32 NHWC copy of every tensor that feeds a residual add, is returned to the caller, or
    feeds the guidance loss (keeps the identity path out of bf16, see DESIGN.md "Tolerances").
  * training only: bf16 raw conv outputs (pre-BN) + per-channel fp64 statistics, fp32 activation gradients.
  * weights: bf16 [tap][cout][cin] packs for forward and (flipped / transposed) for the input gradient, refreshed
    from the fp32 master parameters; eval mode folds BatchNorm into the pack and a bias.

Reference call sites replaced: every nn.Conv2d / ConvTranspose2d / BatchNorm2d / ReLU / ReflectionPad2d /
F.interpolate / torch.cat in /root/reference/src/AE_model_unet.py (see graph.py) and their autograd.
"""
import ctypes as C
import math
import os

import torch

from . import _lib
from ._lib import Act, ConvDesc, WgradDesc
from .graph import Graph, Unit

_SYNC_DEBUG = os.environ.get("GDN_SYNC_DEBUG", "0") == "1"
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


class ActFwdDesc(C.Structure):
    _fields_ = [("src_bf16", C.c_void_p), ("src_f32", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("resid", C.c_void_p), ("relu", C.c_int32), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("c", C.c_int32), ("out_f32", C.c_void_p), ("out_bf16", C.c_void_p), ("pad", C.c_int32),
                ("reflect", C.c_int32), ("up", C.c_int32), ("dilate", C.c_int32), ("src16_is_half", C.c_int32)]


class BnBwdDesc(C.Structure):
    _fields_ = [("dact", C.c_void_p), ("raw", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("mean", C.c_void_p), ("rstd", C.c_void_p), ("relu", C.c_int32), ("n", C.c_int32), ("h", C.c_int32),
                ("w", C.c_int32), ("c", C.c_int32), ("sum_g", C.c_void_p), ("sum_gx", C.c_void_p), ("dy", C.c_void_p),
                ("dilate", C.c_int32), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("raw_is_half", C.c_int32),
                ("dact_is_bf16", C.c_int32)]


class FrozenBwdDesc(C.Structure):
    _fields_ = [("dact", C.c_void_p), ("y_f32", C.c_void_p), ("y_bf16", C.c_void_p), ("scale", C.c_void_p),
                ("relu", C.c_int32), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("dy", C.c_void_p)]


class FoldDesc(C.Structure):
    _fields_ = [("dpad", C.c_void_p), ("ctot", C.c_int32), ("c_off", C.c_int32), ("n", C.c_int32), ("h", C.c_int32),
                ("w", C.c_int32), ("c", C.c_int32), ("pad", C.c_int32), ("reflect", C.c_int32), ("up", C.c_int32),
                ("dilate", C.c_int32), ("dact", C.c_void_p), ("accumulate", C.c_int32), ("dpad_is_bf16", C.c_int32)]


class PackDesc(C.Structure):
    _fields_ = [("kh", C.c_int32), ("kw", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("a_pad", C.c_int32),
                ("b_pad", C.c_int32), ("stride_a", C.c_int64), ("stride_b", C.c_int64), ("stride_r", C.c_int64),
                ("stride_s", C.c_int64), ("flip", C.c_int32), ("col_c", C.c_int32)]


def _ptr(t):
    return None if t is None else t.data_ptr()


def _round_up(v, m):
    return (v + m - 1) // m * m


class _U:
    """per-unit compiled state"""
    pass


class Engine:
    def __init__(self, graph: Graph, params: dict, N: int, H: int, W: int, train: bool, backward: bool = False,
                 want=(), stop_after: str = None, device=None, input_grad: bool = False, grad_seeds=()):
        """params: state_dict-like mapping (no 'module.' prefix) to the module's OWN cuda fp32 tensors
        (parameters are read in place, BN running statistics are updated in place in train mode).
        want: tensor names whose fp32 NHWC value must be available after forward().
        backward with train=False builds the FROZEN backward plan (activation gradients only, BatchNorm folded, no
        parameter gradients): the caller writes dL/d(t) into ``dact[t]`` for every t in grad_seeds, run_backward()
        propagates them (input_grad=True: down to ``dact["in"]``, also for thin 1-channel inputs)."""
        self.L = _lib.lib()
        # GDN_DETERMINISTIC=1 (read once by the library): fixed-order fp32 reductions; the engine's part is the slab
        # workspace of the split-K weight gradients (include/gdn_b200.h, gdn_wgrad_desc.slabs)
        self.det = bool(self.L.gdn_deterministic())
        self.g, self.P, self.N, self.H, self.W = graph, params, N, H, W
        self.train, self.do_bwd = train, backward
        self.frozen_bwd = backward and not train
        self.input_grad, self.grad_seeds = bool(input_grad), tuple(grad_seeds)
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        self.want = set(want)
        units = []
        for u in graph.units:
            units.append(u)
            if stop_after is not None and u.out == stop_after:
                break
        self.units = units
        self.thin_in = graph.cin < 64          # thin inputs go through im2col; wide ones (stand-alone blocks) are tensors
        self._conv_descs = []
        # backward runs on two streams: the weight-gradient chain (wgrad + unpack) is independent of the
        # dgrad -> BatchNorm-backward -> dgrad chain, so its tensor-core kernels overlap the HBM-bound elementwise
        # kernels of the main chain (they co-reside on an SM: the elementwise CTAs need no shared memory)
        self.side_stream = None
        if backward and train and os.environ.get("GDN_SIDE", "1") != "0":
            self.side_stream = torch.cuda.Stream(device=self.dev, priority=-1)
        self._infer_shapes()
        self._plan_tensors()
        self._alloc()
        self._build_forward()
        if backward:
            if train:
                self._build_backward()
            else:
                self._build_backward_frozen()
        self._resolve_first_use()
        if os.environ.get("GDN_PACK_TABLE", "1") != "0":
            self.pack_ops = self._batch_packs(self.pack_ops)
            if backward:
                self.pack_ops_bwd = self._batch_packs(self.pack_ops_bwd)
        self._wversion = None
        self.timeline, self.tl_tag = None, "fwd"
        # (deterministic mode keeps the library heuristics: a variant picked by TIME may differ between two runs, and the
        # staging variants group the pixels differently, i.e. sum the BatchNorm statistics in a different fp32 order)
        if os.environ.get("GDN_AUTOTUNE", "1") != "0" and not self.det and not torch.cuda.is_current_stream_capturing():
            self.autotune()
        self._attach_bn_finalize()

    # ------------------------------------------------------------------ shapes
    def _infer_shapes(self):
        self.shape = {"in": (self.g.cin, self.H, self.W)}
        self.producer = {}
        for u in self.units:
            c0, h, w = self.shape[u.srcs[0]]
            if u.up:
                h, w = 2 * h, 2 * w
            if u.transposed:
                ho = (h - 1) * u.stride - 2 * u.pad + u.k
                wo = (w - 1) * u.stride - 2 * u.pad + u.k
            else:
                ho = (h + 2 * u.pad - u.k) // u.stride + 1
                wo = (w + 2 * u.pad - u.k) // u.stride + 1
            ctot = sum(self.shape[s][0] for s in u.srcs)
            assert ctot == u.cin, (u.conv, ctot, u.cin)
            self.shape[u.out] = (u.cout, ho, wo)
            self.producer[u.out] = u

    @staticmethod
    def _variant(u: Unit):
        """(up, border, reflect, dilate) of the bf16 buffer unit u reads"""
        dil = 1 if (u.transposed and u.stride == 2) else 0
        return (u.up, u.pad if u.reflect else 0, 1 if u.reflect else 0, dil)

    def _plan_tensors(self):
        self.variants = {t: [] for t in self.shape}
        self.need_f32 = {t: (t in self.want) for t in self.shape}
        self.consumers = {t: [] for t in self.shape}
        for u in self.units:
            for s in u.srcs:
                self.consumers[s].append(u)
                if s == "in" and self.thin_in:
                    continue
                v = self._variant(u)
                if v not in self.variants[s]:
                    self.variants[s].append(v)
            if u.resid:
                self.need_f32[u.resid] = True
        if not self.thin_in:
            self.need_f32["in"] = True
        for t, vs in self.variants.items():
            if t == "in":
                continue
            # eval epilogues write plain / reflected borders directly; everything else is derived from the fp32 copy
            if not self.train:
                direct = [v for v in vs if not v[0] and not v[3]]
                if len(direct) != len(vs) or len(direct) > 1:
                    self.need_f32[t] = True
        # network heads have no bf16 consumers; their value is the fp32 output
        for u in self.units:
            if u.tanh:
                self.need_f32[u.out] = True
            # frozen backward reads the ReLU mask from the unit's own output: plain bf16 buffer or the fp32 copy
            if self.frozen_bwd and u.relu and (0, 0, 0, 0) not in self.variants[u.out]:
                self.need_f32[u.out] = True

    # ------------------------------------------------------------------ buffers
    def _buf_dims(self, t, v):
        c, h, w = self.shape[t]
        up, pad, refl, dil = v
        s = 2 if (up or dil) else 1
        return (self.N, h * s + 2 * pad, w * s + 2 * pad, c)

    def _alloc(self):
        dev = self.dev
        self.act = {}    # (tensor, variant) -> bf16 buffer
        self.f32 = {}    # tensor -> fp32 NHWC
        nbytes = 0
        for t, vs in self.variants.items():
            if t == "in" and self.thin_in:
                continue
            for v in vs:
                self.act[(t, v)] = torch.empty(self._buf_dims(t, v), dtype=torch.bfloat16, device=dev)
                nbytes += self.act[(t, v)].numel() * 2
            if self.need_f32[t]:
                c, h, w = self.shape[t]
                self.f32[t] = torch.empty((self.N, h, w, c), dtype=torch.float32, device=dev)
                nbytes += self.f32[t].numel() * 4
        self.activation_bytes = nbytes

    def _act_struct(self, t, v):
        buf = self.act[(t, v)]
        n, hp, wp, c = buf.shape
        pad = v[1]
        return Act(buf.data_ptr(), n, hp - 2 * pad, wp - 2 * pad, c, pad)

    # ------------------------------------------------------------------ helpers to emit calls
    def _call(self, fn, desc, what):
        L = self.L
        if fn is L.gdn_conv2d:
            self._conv_descs.append((what, desc))

        def run(s, fn=fn, desc=desc, what=what):
            rc = fn(C.byref(desc), s)
            if rc:
                _lib.check(rc, what)
            if _SYNC_DEBUG:      # GDN_SYNC_DEBUG=1: attribute asynchronous kernel failures to the op that caused them
                try:
                    torch.cuda.synchronize()
                except Exception as e:
                    raise RuntimeError("gdn_b200: kernel failure in '%s' (algo 0x%x): %s" % (what, getattr(desc, "algo", 0), e))
        run.label = what
        # algorithmic work of tensor-core launches (2*M*N*K with the real channel counts), for tools/profile_ops.py
        if fn is L.gdn_conv2d:
            cin = desc.src0.c + (desc.src1.c if desc.src1.ptr else 0)
            run.flops = 2.0 * desc.src0.n * desc.out_h * desc.out_w * desc.cout * cin * desc.kh * desc.kw
            run.tiles = (desc.src0.n * desc.out_h * desc.out_w + 127) // 128
        elif fn is L.gdn_conv2d_wgrad:
            cin = desc.x0.c + (desc.x1.c if desc.x1.ptr else 0)
            run.flops = 2.0 * desc.x0.n * desc.out_h * desc.out_w * desc.cout_pad * cin * desc.kh * desc.kw
        return run

    def _pack_call(self, pd, wt, scale, out, what, w_off=0):
        L = self.L
        wptr = wt.data_ptr() + 4 * w_off

        def run(s):
            rc = L.gdn_pack_weights(C.byref(pd), C.c_void_p(wptr), C.c_void_p(_ptr(scale)), C.c_void_p(out.data_ptr()), s)
            if rc:
                _lib.check(rc, what)
        run.pack_job = (pd, wptr, _ptr(scale), out.data_ptr(), what)
        run.label = "pack"
        return run

    def _batch_packs(self, ops):
        """replace the per-tensor pack launches in `ops` by table-driven launches (gdn_pack_weights_table), one per
        kernel size (the table kernel's shared-memory tile is sized by the tap count: one table for everything would
        run the many 3x3 layers at the occupancy of a 9x9 tile).  Ops that are not packs (eval-mode BatchNorm folds)
        run first; non-tileable packs (im2col'd thin layers) stay as they are."""
        L = self.L
        jsz = L.gdn_pack_job_size()
        others, left, groups = [], [], {}
        for op in ops:
            job = getattr(op, "pack_job", None)
            if job is None:
                others.append(op)
                continue
            pd, wptr, sptr, optr, what = job
            grp = groups.setdefault(pd.kh * pd.kw, {"blobs": [], "cta0": 0, "first_use": None})
            fu = getattr(op, "first_use", 0)
            grp["first_use"] = fu if grp["first_use"] is None else min(grp["first_use"], fu)
            buf = C.create_string_buffer(jsz)
            n = C.c_int(0)
            rc = L.gdn_pack_job_fill(C.byref(pd), C.c_void_p(wptr), C.c_void_p(sptr), C.c_void_p(optr), grp["cta0"], buf,
                                     C.byref(n))
            if rc:
                left.append(op)
                continue
            grp["blobs"].append(buf.raw)
            grp["cta0"] += n.value
        if sum(len(g["blobs"]) for g in groups.values()) < 2:
            return ops
        runs = []
        for taps, grp in sorted(groups.items()):
            if not grp["blobs"]:
                continue
            table = torch.frombuffer(bytearray(b"".join(grp["blobs"])), dtype=torch.uint8).to(self.dev)

            def run(s, table=table, njobs=len(grp["blobs"]), total=grp["cta0"], taps=taps):
                rc = L.gdn_pack_weights_table(C.c_void_p(table.data_ptr()), njobs, total, taps, s)
                if rc:
                    _lib.check(rc, "pack_weights_table")
            run.label = "pack-table"
            run.first_use = grp["first_use"] or 0
            runs.append(run)
        return others + runs + left

    # ------------------------------------------------------------------ forward plan
    def _geom(self, u: Unit):
        """conv geometry as executed: (k, stride_exec, off, out_h, out_w, flip)"""
        c, ho, wo = self.shape[u.out]
        if u.transposed:
            # ConvTranspose2d == stride-1 conv with the flipped kernel over the (zero-dilated when stride 2) input
            return u.k, 1, -(u.k - 1 - u.pad), ho, wo, 1
        return u.k, u.stride, -u.pad, ho, wo, 0

    def _build_forward(self):
        L, N, dev = self.L, self.N, self.dev
        self.fwd = []          # list of callables(stream)
        self.pack_ops = []     # weight (re)packing, run when parameters changed
        self.cu = {}
        self.launches_fwd = 0
        P = self.P
        if self.train:
            # per-channel sum / sum-of-squares accumulators of EVERY BatchNorm of the network in one fp64 buffer,
            # cleared by one launch at the head of the forward plan (instead of one tiny fill per layer on the
            # critical chain)
            # (+ one 8-byte slot per BatchNorm layer behind the sums: the "statistics flushed" counters of the fused
            # finalisation, cleared by the same launch)
            n_stat = sum(2 * u.cout for u in self.units if u.bn is not None)
            self._fin_slot = n_stat + 2
            self.stat_all = torch.zeros(n_stat + 2 + sum(1 for u in self.units if u.bn is not None), dtype=torch.float64,
                                        device=dev)
            zero_stats = lambda s: self.stat_all.zero_()
            zero_stats.label = "misc"
            self.fwd.append(zero_stats)
            self.launches_fwd += 1
            stat_off = 0
        if not self.thin_in:
            c_in, h_in, w_in = self.shape["in"]
            self.fwd.append(lambda s: self.f32["in"].copy_(self._x.permute(0, 2, 3, 1)))
            for v in self.variants["in"]:
                a = ActFwdDesc()
                a.src_f32 = self.f32["in"].data_ptr()
                a.n, a.h, a.w, a.c = N, h_in, w_in, c_in
                a.out_bf16 = self.act[("in", v)].data_ptr()
                a.up, a.pad, a.reflect, a.dilate = v
                self.fwd.append(self._call(L.gdn_act_forward, a, "input variant"))
                self.launches_fwd += 1
        for u in self.units:
            cu = _U()
            self.cu[u.conv] = cu
            cu.u = u
            k, stride, off, ho, wo, flip = self._geom(u)
            cu.k, cu.stride, cu.off, cu.ho, cu.wo = k, stride, off, ho, wo
            wt = P[u.conv + ".weight"]
            cout_pad = 16 if u.cout < 16 else u.cout
            cu.cout_pad = cout_pad
            thin = u.cin < 64
            cu.thin = thin
            kk = k * k
            if (u.cout == 1 and not thin and k == 9 and stride == 1 and len(u.srcs) == 1 and not u.up and u.bn is None
                    and not u.reflect and u.resid is None and u.cin % 64 == 0 and os.environ.get("GDN_HEAD_TAPS", "1") != "0"):
                self._build_head_forward(u, cu, wt, flip)
                continue
            # ---------------- weights (forward pack)
            if thin:
                kpad = _round_up(kk * u.cin, 64)
                cu.kpad = kpad
                cu.wf = torch.empty((1, cout_pad, kpad), dtype=torch.bfloat16, device=dev)
                pd = PackDesc(k, k, u.cout, kpad, cout_pad, kpad, u.cin * kk, kk, k, 1, 0, u.cin)
                assert not u.transposed
            else:
                cu.wf = torch.empty((kk, cout_pad, u.cin), dtype=torch.bfloat16, device=dev)
                if u.transposed:
                    pd = PackDesc(k, k, u.cout, u.cin, cout_pad, u.cin, kk, u.cout * kk, k, 1, 1, 0)
                else:
                    pd = PackDesc(k, k, u.cout, u.cin, cout_pad, u.cin, u.cin * kk, kk, k, 1, 0, 0)
            cu.pd_fwd = pd
            bn_eval = (u.bn is not None) and not self.train
            if bn_eval:
                cu.fold_scale = torch.empty(u.cout, dtype=torch.float32, device=dev)
                cu.fold_bias = torch.empty(u.cout, dtype=torch.float32, device=dev)
                g_, b_ = P[u.bn + ".weight"], P[u.bn + ".bias"]
                rm, rv = P[u.bn + ".running_mean"], P[u.bn + ".running_var"]

                def fold(s, g_=g_, b_=b_, rm=rm, rv=rv, cu=cu, c=u.cout):
                    rc = L.gdn_bn_fold(C.c_void_p(g_.data_ptr()), C.c_void_p(b_.data_ptr()), C.c_void_p(rm.data_ptr()),
                                       C.c_void_p(rv.data_ptr()), C.c_float(BN_EPS), C.c_void_p(cu.fold_scale.data_ptr()),
                                       C.c_void_p(cu.fold_bias.data_ptr()), c, s)
                    if rc:
                        _lib.check(rc, "bn_fold")
                fold.cu = cu
                self.pack_ops.append(fold)
                self.pack_ops.append(self._pack_call(pd, wt, cu.fold_scale, cu.wf, "pack " + u.conv))
            else:
                self.pack_ops.append(self._pack_call(pd, wt, None, cu.wf, "pack " + u.conv))
            self.pack_ops[-1].cu = cu

            # ---------------- input operand(s)
            d = ConvDesc()
            if thin:
                src = u.srcs[0]
                assert len(u.srcs) == 1 and not u.up and stride == 1
                c_in, h_in, w_in = self.shape[src]
                cu.col = torch.empty((N, h_in, w_in, cu.kpad), dtype=torch.bfloat16, device=dev)
                cu.col_src = src

                def im2col(s, cu=cu, u=u, h_in=h_in, w_in=w_in, k=k):
                    x = self._thin_input(cu.col_src)
                    rc = L.gdn_im2col(C.c_void_p(x.data_ptr()), C.c_void_p(cu.col.data_ptr()), N, u.cin, h_in, w_in, k, k,
                                      u.pad, 1 if u.reflect else 0, cu.kpad, s)
                    if rc:
                        _lib.check(rc, "im2col " + u.conv)
                self.fwd.append(im2col)
                d.src0 = Act(cu.col.data_ptr(), N, h_in, w_in, cu.kpad, 0)
                d.kh = d.kw = 1
                d.stride = 1
                d.off_y = d.off_x = 0
            else:
                v = self._variant(u)
                d.src0 = self._act_struct(u.srcs[0], v)
                if len(u.srcs) == 2:
                    d.src1 = self._act_struct(u.srcs[1], v)
                d.kh = d.kw = k
                d.stride = stride
                d.off_y = d.off_x = off
            d.weights = cu.wf.data_ptr()
            d.out_h, d.out_w = ho, wo
            d.cout, d.cout_pad = u.cout, cout_pad
            d.algo = 0
            d.dst_h, d.dst_w = ho, wo
            d.dst_sy = d.dst_sx = 1
            cu.conv_desc = d
            out_vs = self.variants[u.out]
            if self.train and u.bn is not None:
                # raw conv output + statistics, then BatchNorm apply into every variant the consumers need
                cu.raw = torch.empty((N, ho, wo, u.cout), dtype=torch.float16, device=dev)
                cu.stat = self.stat_all[stat_off:stat_off + 2 * u.cout].view(2, u.cout)
                stat_off += 2 * u.cout
                cu.scale = torch.empty(u.cout, dtype=torch.float32, device=dev)
                cu.shift = torch.empty(u.cout, dtype=torch.float32, device=dev)
                cu.mean = torch.empty(u.cout, dtype=torch.float32, device=dev)
                cu.rstd = torch.empty(u.cout, dtype=torch.float32, device=dev)
                cu.coef4 = torch.empty((u.cout, 4), dtype=torch.float32, device=dev)   # (scale, shift, mean, rstd) interleaved
                d.out_bf16 = Act(cu.raw.data_ptr(), N, ho, wo, u.cout, 0)
                d.out16_is_half = 1
                d.stat_sum = cu.stat[0].data_ptr()
                d.stat_sqsum = cu.stat[1].data_ptr()
                cu.fwd_conv_idx = len(self.fwd)
                self.fwd.append(self._call(L.gdn_conv2d, d, "conv " + u.conv))
                g_, b_ = P[u.bn + ".weight"], P[u.bn + ".bias"]
                rm, rv = P[u.bn + ".running_mean"], P[u.bn + ".running_var"]
                cnt = float(N * ho * wo)

                # BatchNorm finalisation (statistics -> scale / shift / mean / rstd, running statistics): by the LAST CTA of
                # the convolution itself (gdn_conv_desc.fin_*), attached after the autotuner has run (its timing launches
                # must not touch the running statistics); the separate launch remains for GDN_FUSE_BNFIN=0
                cu.fin = (self.stat_all[self._fin_slot:self._fin_slot + 1], g_, b_, rm, rv, cnt)
                self._fin_slot += 1
                cu.fin_fused = False

                def finalize(s, cu=cu, g_=g_, b_=b_, rm=rm, rv=rv, cnt=cnt, c=u.cout):
                    if cu.fin_fused:
                        return
                    rc = L.gdn_bn_finalize(C.c_void_p(cu.stat[0].data_ptr()), C.c_void_p(cu.stat[1].data_ptr()),
                                           C.c_double(cnt), C.c_void_p(g_.data_ptr()), C.c_void_p(b_.data_ptr()),
                                           C.c_float(BN_EPS), C.c_float(BN_MOMENTUM), C.c_void_p(rm.data_ptr()),
                                           C.c_void_p(rv.data_ptr()), C.c_void_p(cu.scale.data_ptr()),
                                           C.c_void_p(cu.shift.data_ptr()), C.c_void_p(cu.mean.data_ptr()),
                                           C.c_void_p(cu.rstd.data_ptr()), C.c_void_p(cu.coef4.data_ptr()), c, s)
                    if rc:
                        _lib.check(rc, "bn_finalize " + u.conv)
                self.fwd.append(finalize)
                self.launches_fwd += 2
                first = True
                todo = list(out_vs) if out_vs else [None]
                for v in todo:
                    a = ActFwdDesc()
                    a.src_bf16 = cu.raw.data_ptr()
                    a.src16_is_half = 1
                    a.scale, a.shift = cu.scale.data_ptr(), cu.shift.data_ptr()
                    a.resid = _ptr(self.f32.get(u.resid)) if u.resid else None
                    a.relu = int(u.relu)
                    a.n, a.h, a.w, a.c = N, ho, wo, u.cout
                    if first and self.need_f32[u.out]:
                        a.out_f32 = self.f32[u.out].data_ptr()
                    if v is not None and v[0]:
                        # x2 bilinear variant: BatchNorm / ReLU / residual ONCE at low resolution into a plain bf16
                        # buffer, then a pure interpolation pass over it (4 cached 16-byte taps per output vector)
                        # instead of normalising every tap of every output again
                        plain = (0, 0, 0, 0)
                        if (u.out, plain) not in self.act:
                            self.act[(u.out, plain)] = torch.empty((N, ho, wo, u.cout), dtype=torch.bfloat16, device=dev)
                        a.out_bf16 = self.act[(u.out, plain)].data_ptr()
                        self.fwd.append(self._call(L.gdn_act_forward, a, "act " + u.conv))
                        b = ActFwdDesc()
                        b.src_bf16 = self.act[(u.out, plain)].data_ptr()
                        b.n, b.h, b.w, b.c = N, ho, wo, u.cout
                        b.out_bf16 = self.act[(u.out, v)].data_ptr()
                        b.up, b.pad, b.reflect, b.dilate = v
                        self.fwd.append(self._call(L.gdn_act_forward, b, "act-up " + u.conv))
                        self.launches_fwd += 2
                        first = False
                        continue
                    if v is not None:
                        a.out_bf16 = self.act[(u.out, v)].data_ptr()
                        a.up, a.pad, a.reflect, a.dilate = v
                    if a.out_f32 or a.out_bf16:
                        self.fwd.append(self._call(L.gdn_act_forward, a, "act " + u.conv))
                        self.launches_fwd += 1
                    first = False
            else:
                # eval (BN folded) or no BN: the conv epilogue produces the activation itself
                if u.bn is not None:
                    d.bias = cu.fold_bias.data_ptr()
                d.relu = int(u.relu)
                d.tanh_out = int(u.tanh)
                if u.resid:
                    d.resid = self.f32[u.resid].data_ptr()
                if self.need_f32[u.out]:
                    d.out_f32 = self.f32[u.out].data_ptr()
                direct = [v for v in out_vs if not v[0] and not v[3]]
                derived = [v for v in out_vs if v[0] or v[3]]
                if direct:
                    v = direct[0]
                    d.out_bf16 = self._act_struct(u.out, v)
                    d.out_reflect = v[2]
                cu.fwd_conv_idx = len(self.fwd)
                self.fwd.append(self._call(L.gdn_conv2d, d, "conv " + u.conv))
                self.launches_fwd += 1
                for v in direct[1:] + derived:
                    a = ActFwdDesc()
                    a.src_f32 = self.f32[u.out].data_ptr()
                    a.n, a.h, a.w, a.c = N, ho, wo, u.cout
                    a.out_bf16 = self.act[(u.out, v)].data_ptr()
                    a.up, a.pad, a.reflect, a.dilate = v
                    self.fwd.append(self._call(L.gdn_act_forward, a, "variant " + u.conv))
                    self.launches_fwd += 1
                if (direct[1:] or derived) and not self.need_f32[u.out]:
                    raise AssertionError("variant derivation needs the fp32 copy of " + u.out)

    def _build_head_forward(self, u, cu, wt, flip):
        """64 -> 1 head (k9, zero pad 4, tanh; AE_model_unet.py:300,362 / :521,570).  As a convolution it has ONE output
        channel: a 16-wide MMA tile whose 81 taps each re-read the activation window from shared memory (0.53 ms at
        21 TFLOP/s, profiles/r02f_profile_ops.log).  Here the taps are the N dimension of one 1x1 convolution,
        Z[p][t] = <x[p], w[t]> (fp16, 128 columns), and gdn_head_gather sums Z over the 9x9 neighbourhood and applies tanh."""
        L, N, dev = self.L, self.N, self.dev
        k, kk, ho, wo = cu.k, cu.k * cu.k, cu.ho, cu.wo
        zc = _round_up(kk, 64)
        cu.cout_pad = zc
        cu.wf = torch.empty((1, zc, u.cin), dtype=torch.bfloat16, device=dev)
        # W[t][c]: Conv2d weight [1][c][r][s] and ConvTranspose2d weight [c][1][r][s] are both c*kk + tap in memory; the
        # transposed head is a convolution with the flipped kernel (tap kk-1-t)
        if flip:
            pd, w_off = PackDesc(1, 1, kk, u.cin, zc, u.cin, -1, kk, 0, 0, 0, 0), kk - 1
        else:
            pd, w_off = PackDesc(1, 1, kk, u.cin, zc, u.cin, 1, kk, 0, 0, 0, 0), 0
        cu.pd_fwd = pd
        op = self._pack_call(pd, wt, None, cu.wf, "pack " + u.conv, w_off=w_off)
        op.cu = cu
        self.pack_ops.append(op)
        cu.z = torch.empty((N, ho, wo, zc), dtype=torch.float16, device=dev)
        d = ConvDesc()
        d.src0 = self._act_struct(u.srcs[0], self._variant(u))
        d.kh = d.kw = 1
        d.stride = 1
        d.off_y = d.off_x = 0
        d.weights = cu.wf.data_ptr()
        d.out_h, d.out_w = ho, wo
        d.cout = d.cout_pad = zc
        d.algo = 0
        d.dst_h, d.dst_w = ho, wo
        d.dst_sy = d.dst_sx = 1
        d.out_bf16 = Act(cu.z.data_ptr(), N, ho, wo, zc, 0)
        d.out16_is_half = 1
        cu.conv_desc = d
        cu.fwd_conv_idx = len(self.fwd)
        self.fwd.append(self._call(L.gdn_conv2d, d, "conv " + u.conv))
        out = self.f32[u.out]
        pad_exec, tanh = -cu.off, int(u.tanh)

        def gather(s, cu=cu, out=out):
            rc = L.gdn_head_gather(C.c_void_p(cu.z.data_ptr()), 1, zc, N, ho, wo, k, pad_exec, tanh,
                                   C.c_void_p(out.data_ptr()), s)
            if rc:
                _lib.check(rc, "head_gather " + u.conv)
        gather.label = "head-gather " + u.conv
        self.fwd.append(gather)
        self.launches_fwd += 2

    def _resolve_first_use(self):
        """index of the forward op that first reads each weight pack (for packs issued on the side stream)"""
        for op in self.pack_ops:
            cu = getattr(op, "cu", None)
            op.first_use = getattr(cu, "fwd_conv_idx", 0) if cu is not None else 0

    def _thin_input(self, name):
        if name == "in":
            return self._x
        return self.f32[name]

    # ------------------------------------------------------------------ running
    def refresh_weights(self, stream=None):
        s = stream or _lib.stream_ptr()
        side = self.side_stream
        self._fwd_wait = None
        if (side is not None and self.train and getattr(self, "async_fwd_pack", False)
                and os.environ.get("GDN_ASYNC_FWD_PACK", "0") == "1"):
            # EXPERIMENT, off by default: the forward packs run on the side stream in order of first use and the
            # forward plan waits for each pack just before the first convolution that reads it, so that the big
            # 3x3 / 512-channel tables are re-packed UNDER the first layers instead of in front of them (~0.75 ms at the
            # head of the step).  Measured on the B200 (profiles/r01r_bench_{async,sync}pack.json, same box, back to
            # back): 535.7 img/s with, 543.5 without -- the 28 k small pack CTAs delay the persistent conv CTAs of the
            # first layers by more than the serial re-pack costs.  Kept for the next round (a faster pack kernel).
            main = torch.cuda.current_stream(self.dev)
            side.wait_stream(main)
            waits = {}
            with torch.cuda.stream(side):
                sp = C.c_void_p(side.cuda_stream)
                for op in sorted(self.pack_ops, key=lambda o: getattr(o, "first_use", 0)):
                    op(sp)
                    ev = torch.cuda.Event()
                    ev.record(side)
                    waits[getattr(op, "first_use", 0)] = ev     # same stream: a later event covers the earlier packs
            self._fwd_wait = waits
        else:
            for op in self.pack_ops:
                op(s)
        if self.do_bwd:
            side = self.side_stream
            if side is not None and getattr(self, "async_bwd_pack", False):
                # the dgrad weight packs are only needed by backward: re-pack them on the side stream, under the
                # forward convolutions (fused training steps set async_bwd_pack; they always run backward next)
                main = torch.cuda.current_stream(self.dev)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    for op in self.pack_ops_bwd:
                        op(C.c_void_p(side.cuda_stream))
                    self._pack_ev = torch.cuda.Event()
                    self._pack_ev.record(side)
            else:
                for op in self.pack_ops_bwd:
                    op(s)

    def refresh_if_stale(self):
        """re-pack the weights now if the parameters changed since the last pack (forward() then finds them fresh)"""
        ver = self._param_version()
        if ver != self._wversion:
            self.refresh_weights(_lib.stream_ptr())
            self._wversion = ver

    def _param_version(self):
        return tuple(t._version for t in self.P.values())

    def forward(self, x):
        """x: fp32 NCHW cuda tensor (N, cin, H, W).  Returns nothing; read results with value()/depth()."""
        if x.shape != (self.N, self.g.cin, self.H, self.W):
            raise ValueError("gdn_b200 engine built for %s, got %s" % ((self.N, self.g.cin, self.H, self.W), tuple(x.shape)))
        if x.dtype != torch.float32 or not x.is_cuda:
            raise ValueError("gdn_b200: input must be a CUDA fp32 tensor")
        self._x = x.contiguous()
        s = _lib.stream_ptr()
        ver = self._param_version()
        if ver != self._wversion:
            self.refresh_weights(s)
            self._wversion = ver
        tl = getattr(self, "timeline", None)
        waits, self._fwd_wait = getattr(self, "_fwd_wait", None), None
        if waits:
            main = torch.cuda.current_stream(self.dev)
            pending = sorted(waits.items())

            def gate(i):                    # packs issued on the side stream: wait right before their first reader
                while pending and pending[0][0] <= i:
                    main.wait_event(pending.pop(0)[1])
        else:
            gate = None
        if tl is None:
            for i, op in enumerate(self.fwd):
                if gate is not None:
                    gate(i)
                op(s)
        else:                               # development aid (tools/timeline.py): completion event after every op
            cur = torch.cuda.current_stream(self.dev)
            for i, op in enumerate(self.fwd):
                if gate is not None:
                    gate(i)
                op(s)
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(cur)
                tl.append((getattr(op, "label", "misc"), self.tl_tag, ev))
        if self.train:
            self._wversion = None if self.do_bwd else self._param_version()  # running stats were updated in place

    def value(self, name):
        """fp32 NHWC buffer of a tensor (valid until the next forward)"""
        return self.f32[name]

    def value_nchw(self, name):
        """(N, C, H, W)-shaped view (channels_last strides) of the fp32 buffer"""
        return self.f32[name].permute(0, 3, 1, 2)

    def depth(self):
        """network output (N, 1, H, W) fp32 (C == 1, so NHWC and NCHW coincide)"""
        t = self.units[-1].out
        return self.f32[t].view(self.N, 1, self.H, self.W)

    # ------------------------------------------------------------------ backward plan
    def _build_backward(self):
        L, N, dev, P = self.L, self.N, self.dev, self.P
        assert self.train
        self.bwd = []
        self.pack_ops_bwd = []
        self.grad_ready_op = {}   # parameter name -> index of the backward op after which its gradient is final
        self.launches_bwd = 0
        # parameter gradients: one flat fp32 buffer, views per parameter in state_dict order
        names = [k for k, t in P.items() if t.dtype == torch.float32 and getattr(t, "requires_grad", False)]
        self.param_names = names
        total = sum(_round_up(P[k].numel(), 4) for k in names)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = {}
        o = 0
        for k in names:
            n = P[k].numel()
            self.grad[k] = self.flat_grad[o:o + n].view(P[k].shape)
            o += _round_up(n, 4)
        # activation gradients (fp32 NHWC) for every tensor except the input
        self.dact = {}
        for t in self.shape:
            if t == "in" and self.thin_in:
                continue
            c, h, w = self.shape[t]
            self.dact[t] = torch.empty((N, h, w, c), dtype=torch.float32, device=dev)
        self._dset = {}   # tensor -> bool (runtime: has a gradient been written in this backward pass?)
        self._pending_add = {}   # tensor -> fp32 buffer still to be added to its gradient by the next dgrad conv
        maxdw = 0
        for u in self.units:
            cu = self.cu[u.conv]
            maxdw = max(maxdw, (1 if cu.thin else cu.k * cu.k) * (cu.kpad if cu.thin else u.cin) * max(cu.cout_pad, 64))
        maxdw = max(maxdw, 128 * 64)
        self.dw_scratch = torch.zeros(maxdw, dtype=torch.float32, device=dev)
        # sum(g) / sum(g * xhat) accumulators of every BatchNorm-backward reduction, cleared by ONE launch at the head
        # of the backward plan
        n_bn = sum(1 for u in self.units if u.bn is not None)
        self.bn_sums_all = torch.zeros((max(n_bn, 1), 2, 512), dtype=torch.float64, device=dev)
        zero_sums = lambda s: self.bn_sums_all.zero_()
        zero_sums.label = "misc"
        self.bwd.append(zero_sums)
        self.launches_bwd += 1
        self._bn_idx = {}
        for u in self.units:
            if u.bn is not None:
                self._bn_idx[u.conv] = len(self._bn_idx)
        # Fused BatchNorm-backward statistics: the input-gradient convolution that writes the LAST contribution to a
        # tensor's gradient also reduces sum g / sum g*xhat of the unit that produced the tensor in its epilogue (and
        # applies that unit's ReLU mask), so the separate reduce pass over the gradient disappears; when nothing else
        # reads the fp32 gradient it is written as bf16 only (``gm``), which is all the apply pass needs.
        # last_contrib[t]: the unit whose backward adds the final term to d(t) -- the live consumer / residual user
        # that comes FIRST in forward order (backward visits units in reverse).
        self._fuse_bnbwd = os.environ.get("GDN_FUSE_BNBWD", "1") != "0"
        live = {self.units[-1].out}
        for u in reversed(self.units):
            if u.out in live:
                live.update(u.srcs)
                if u.resid:
                    live.add(u.resid)
        self._last_contrib = {}
        for u in self.units:
            if u.out not in live:
                continue
            for t in tuple(u.srcs) + ((u.resid,) if u.resid else ()):
                self._last_contrib.setdefault(t, u)
        self.gm = {}             # tensor -> bf16 masked total gradient written by the fusing epilogue
        self._fused = set()      # tensors whose producer's BatchNorm-backward sums come from a dgrad epilogue

        order = []  # static accumulation bookkeeping: which tensors already hold a gradient at each point
        have = set()
        have.add(self.units[-1].out)

        def accumulate_flag(t):
            f = t in have
            have.add(t)
            return f

        for u in reversed(self.units):
            cu = self.cu[u.conv]
            k, stride, off, ho, wo = cu.k, cu.stride, cu.off, cu.ho, cu.wo
            kk = k * k
            wt = P[u.conv + ".weight"]
            if u.out not in have:
                continue  # tensor does not influence the loss (e.g. dead branch): no gradient flows
            g_out = self.dact[u.out]
            if u.tanh:
                self._build_head_backward(u, cu, have)
                continue
            # ---- residual identity: d(resid) (+)= g_out
            if u.resid and u.resid not in have and self._next_grad_is_direct_conv(u):
                # nothing has been written to d(resid) yet and the next contribution is a plain dgrad conv:
                # let that conv's epilogue add g_out (its `resid` operand) instead of copying g_out now
                self._pending_add[u.resid] = g_out
                have.add(u.resid)
            elif u.resid:
                acc = accumulate_flag(u.resid)
                a = ActFwdDesc()
                a.src_f32 = g_out.data_ptr()
                a.resid = self.dact[u.resid].data_ptr() if acc else None
                a.n, a.h, a.w, a.c = N, ho, wo, u.cout
                a.out_f32 = self.dact[u.resid].data_ptr()
                self.bwd.append(self._call(L.gdn_act_forward, a, "resid-grad " + u.conv))
                self.launches_bwd += 1
            # ---- BatchNorm (+ReLU) backward -> dy (bf16)
            assert u.bn is not None, "training needs BatchNorm on every hidden conv (" + u.conv + ")"
            cu.dy = torch.empty((N, ho, wo, u.cout), dtype=torch.bfloat16, device=dev)
            b = BnBwdDesc()
            b.dact, b.raw = g_out.data_ptr(), cu.raw.data_ptr()
            b.scale, b.shift, b.mean, b.rstd = (cu.scale.data_ptr(), cu.shift.data_ptr(), cu.mean.data_ptr(),
                                                cu.rstd.data_ptr())
            b.relu = int(u.relu)
            b.raw_is_half = 1
            b.n, b.h, b.w, b.c = N, ho, wo, u.cout
            bn_i = self._bn_idx[u.conv]
            b.sum_g, b.sum_gx = self.bn_sums_all[bn_i, 0].data_ptr(), self.bn_sums_all[bn_i, 1].data_ptr()
            b.dy = cu.dy.data_ptr()
            b.dilate = 0
            b.dgamma = self.grad[u.bn + ".weight"].data_ptr()
            b.dbeta = self.grad[u.bn + ".bias"].data_ptr()
            if u.out in self._fused:
                # the sums were reduced (and the ReLU mask applied) by the epilogue of the convolution that completed g_out
                if u.out in self.gm:
                    b.dact, b.dact_is_bf16, b.relu = self.gm[u.out].data_ptr(), 1, 0
            else:
                self.bwd.append(self._call(L.gdn_bn_bwd_reduce, b, "bn_bwd_reduce " + u.conv))
                self.launches_bwd += 1
            self.bwd.append(self._call(L.gdn_act_backward, b, "act_backward " + u.conv))
            self.grad_ready_op[u.bn + ".weight"] = self.grad_ready_op[u.bn + ".bias"] = len(self.bwd) - 1
            self.launches_bwd += 1
            # ---- weight gradient (same geometry as the forward conv)
            wd = WgradDesc()
            fd = cu.conv_desc
            wd.x0, wd.x1 = fd.src0, fd.src1
            cpad64 = _round_up(cu.cout_pad, 64)
            assert cpad64 == u.cout, "hidden convs have >= 64 output channels"
            wd.dy = Act(cu.dy.data_ptr(), N, ho, wo, u.cout, 0)
            cin_eff = cu.kpad if cu.thin else u.cin
            taps_eff = 1 if cu.thin else kk
            dw = self.dw_scratch[: taps_eff * cin_eff * u.cout]
            wd.dw = dw.data_ptr()
            wd.kh, wd.kw, wd.stride = fd.kh, fd.kw, fd.stride
            wd.off_y, wd.off_x = fd.off_y, fd.off_x
            wd.out_h, wd.out_w, wd.cout_pad = ho, wo, u.cout
            used = self._det_slabs(wd)
            # (the accumulating kernel needs a clean scratch.  Letting the unpack write zeros back to what it read instead
            # of these 46 fills was measured and rejected: unpack 0.77 -> 1.34 ms per step against 0.19 ms of fills,
            # profiles/r02p_profile_ops.log; the slab mode stores and needs neither)
            zero_dw = lambda s, dw=dw: dw.zero_()
            zero_dw.label = "misc"
            side_ops = ([] if used is not None else [zero_dw]) + [self._call(L.gdn_conv2d_wgrad, wd, "wgrad " + u.conv)]
            if cu.thin:
                up_ = PackDesc(k, k, u.cout, cu.kpad, u.cout, cu.kpad, u.cin * kk, kk, k, 1, 0, u.cin)
            elif u.transposed:
                up_ = PackDesc(k, k, u.cout, u.cin, u.cout, u.cin, kk, u.cout * kk, k, 1, 1, 0)
            else:
                up_ = PackDesc(k, k, u.cout, u.cin, u.cout, u.cin, u.cin * kk, kk, k, 1, 0, 0)
            gbuf = self.grad[u.conv + ".weight"]

            def unpack(s, up_=up_, dw=dw, gbuf=gbuf, name=u.conv, used=used):
                if used is None:
                    rc = L.gdn_unpack_wgrad(C.byref(up_), C.c_void_p(dw.data_ptr()), C.c_void_p(gbuf.data_ptr()), 1, s)
                else:       # deterministic split-K: sum the slabs the wgrad launch just filled, in index order
                    rc = L.gdn_unpack_wgrad_slabs(C.byref(up_), C.c_void_p(self.dw_slabs.data_ptr()), used.value,
                                                  C.c_int64(dw.numel()), C.c_void_p(gbuf.data_ptr()), 1, s)
                if rc:
                    _lib.check(rc, "unpack " + name)
            unpack.label = "unpack " + u.conv
            side_ops.append(unpack)
            self.launches_bwd += 2
            # ---- input gradient(s)
            c_off = 0
            self._dgrad_conv_end = len(self.bwd)
            for si, s_name in enumerate(u.srcs):
                cs = self.shape[s_name][0]
                if s_name == "in" and self.thin_in:
                    c_off += cs
                    continue
                self._build_dgrad(u, cu, s_name, c_off, cs, accumulate_flag(s_name))
                c_off += cs
            # the weight-gradient group goes right after this unit's dgrad convolution(s): on the side stream it
            # starts when they have finished and overlaps the fold / BatchNorm-backward kernels that follow
            pos = self._dgrad_conv_end
            for op in side_ops:
                op.side = 1
            self.bwd[pos:pos] = side_ops
            self.grad_ready_op[u.conv + ".weight"] = pos + len(side_ops) - 1

    def _build_backward_frozen(self):
        """Backward plan of an eval-mode (frozen) network: activation gradients only.  Per unit, in reverse order:
        residual identity, dy = g * [y > 0] * folded-BatchNorm scale (gdn_act_backward_frozen), then the same
        input-gradient convolutions as in training, on the UN-folded weight packs (the scale went into dy).
        Used by the opt-in guidance gradient (trainer.RtoDTrainStep(guidance_grad=True)): the
        latent loss back-propagates through the frozen DtoD encoder into the RtoD output; the published
        trainer.py:699-703 blocks that path with no_grad."""
        L, N, dev = self.L, self.N, self.dev
        self.bwd, self.pack_ops_bwd, self.grad_ready_op = [], [], {}
        self.launches_bwd = 0
        self.dact = {}
        for t in self.shape:
            if t == "in" and not self.input_grad:
                continue
            c, h, w = self.shape[t]
            self.dact[t] = torch.zeros((N, h, w, c), dtype=torch.float32, device=dev)
        self._pending_add = {}
        have = set(self.grad_seeds)
        for t in have:
            if t not in self.shape:
                raise ValueError("gdn_b200: gradient seed %r is not a tensor of this graph" % t)

        def accumulate_flag(t):
            f = t in have
            have.add(t)
            return f

        for u in reversed(self.units):
            cu = self.cu[u.conv]
            if u.out not in have:
                continue
            if u.tanh or u.up or (u.relu and u.resid):
                raise NotImplementedError("gdn_b200: frozen backward covers encoder-style units only (%s)" % u.conv)
            ho, wo = cu.ho, cu.wo
            g_out = self.dact[u.out]
            if u.resid and u.resid not in have and self._next_grad_is_direct_conv(u):
                self._pending_add[u.resid] = g_out
                have.add(u.resid)
            elif u.resid:
                acc = accumulate_flag(u.resid)
                a = ActFwdDesc()
                a.src_f32 = g_out.data_ptr()
                a.resid = self.dact[u.resid].data_ptr() if acc else None
                a.n, a.h, a.w, a.c = N, ho, wo, u.cout
                a.out_f32 = self.dact[u.resid].data_ptr()
                self.bwd.append(self._call(L.gdn_act_forward, a, "resid-grad " + u.conv))
                self.launches_bwd += 1
            cu.dy = torch.empty((N, ho, wo, u.cout), dtype=torch.bfloat16, device=dev)
            fb = FrozenBwdDesc()
            fb.dact = g_out.data_ptr()
            fb.relu = int(u.relu)
            if u.relu:
                if u.out in self.f32:
                    fb.y_f32 = self.f32[u.out].data_ptr()
                else:
                    fb.y_bf16 = self.act[(u.out, (0, 0, 0, 0))].data_ptr()
            fb.scale = cu.fold_scale.data_ptr() if u.bn is not None else None
            fb.n, fb.h, fb.w, fb.c = N, ho, wo, u.cout
            fb.dy = cu.dy.data_ptr()
            self.bwd.append(self._call(L.gdn_act_backward_frozen, fb, "act_backward_frozen " + u.conv))
            self.launches_bwd += 1
            c_off = 0
            self._dgrad_conv_end = len(self.bwd)
            for s_name in u.srcs:
                cs = self.shape[s_name][0]
                if s_name == "in" and not self.input_grad:
                    c_off += cs
                    continue
                self._build_dgrad(u, cu, s_name, c_off, cs, accumulate_flag(s_name))
                c_off += cs

    def _next_grad_is_direct_conv(self, u):
        """True when the consumer of u.resid that runs next in backward order (the one closest before u in forward
        order) is a single-source, stride-1, zero-padded convolution -- its dgrad writes d(resid) directly."""
        idx = self.units.index(u)
        for v in reversed(self.units[:idx]):
            if u.resid in v.srcs:
                return (len(v.srcs) == 1 and v.stride == 1 and not v.transposed and not v.up and not v.reflect
                        and v.cin >= 64 and not v.tanh)
        return False

    def _build_dgrad(self, u, cu, s_name, c_off, cs, acc):
        """gradient of unit u w.r.t. source tensor s_name (channels [c_off, c_off+cs) of its input)"""
        L, N, dev, P = self.L, self.N, self.dev, self.P
        k, kk = cu.k, cu.k * cu.k
        wt = P[u.conv + ".weight"]
        v = self._variant(u)
        up, pad_phys, refl, dil = v
        c_s, h_s, w_s = self.shape[s_name]
        # dgrad weight pack: [tap][a = ci of this source][b = co]
        if (not u.transposed) and u.stride == 2:
            self._build_dgrad_stride2(u, cu, s_name, c_off, cs, acc)
            return
        cs_pad = 16 if cs < 16 else cs      # thin network input (frozen backward with input_grad): 16-wide MMA tile
        wdg = torch.empty((kk, cs_pad, u.cout), dtype=torch.bfloat16, device=dev)
        if u.transposed:
            # w is (cin, cout, k, k): dX = conv(dy, w) -- no flip
            pd = PackDesc(k, k, cs, u.cout, cs_pad, u.cout, u.cout * kk, kk, k, 1, 0, 0)
            w_off = c_off * u.cout * kk
        else:
            pd = PackDesc(k, k, cs, u.cout, cs_pad, u.cout, kk, u.cin * kk, k, 1, 1, 0)
            w_off = c_off * kk
        self.pack_ops_bwd.append(self._pack_call(pd, wt, None, wdg, "pack-dgrad " + u.conv, w_off))
        d = ConvDesc()
        d.weights = wdg.data_ptr()
        d.kh = d.kw = k
        d.cout, d.cout_pad = cs, cs_pad
        d.algo = 0
        d.dst_sy = d.dst_sx = 1
        direct = (not up) and (not refl) and (not dil)
        if u.transposed and u.stride == 2:
            # dX[y] = sum_t dy[2y + t - p] w[t]: stride-2 conv over dy, straight onto the source gradient
            d.src0 = Act(cu.dy.data_ptr(), N, cu.ho, cu.wo, u.cout, 0)
            d.stride = 2
            d.off_y = d.off_x = -u.pad
            d.out_h, d.out_w = h_s, w_s
            direct = True
        else:
            offb = cu.off + pad_phys            # forward offset in buffer coordinates
            d.src0 = Act(cu.dy.data_ptr(), N, cu.ho, cu.wo, u.cout, 0)
            d.stride = 1
            d.off_y = d.off_x = -(k - 1) - offb
            sc = 2 if (up or dil) else 1
            d.out_h, d.out_w = h_s * sc + 2 * pad_phys, w_s * sc + 2 * pad_phys
        d.dst_h, d.dst_w = d.out_h, d.out_w
        if direct:
            tgt = self.dact[s_name]
            d.out_f32 = tgt.data_ptr()
            pend = self._pending_add.pop(s_name, None)
            if pend is not None:
                d.resid = pend.data_ptr()
            else:
                d.resid = tgt.data_ptr() if acc else None
            self._fuse_bn_backward_stats(d, u, s_name)
            self.bwd.append(self._call(L.gdn_conv2d, d, "dgrad " + u.conv))
            self._dgrad_conv_end = len(self.bwd)
            self.launches_bwd += 1
        else:
            # gradient w.r.t. the padded / upsampled / dilated operand buffer: bf16 (half the bytes of the largest
            # intermediate of the backward pass; the fold sums <= 36 of them per source element in fp32)
            tmp16 = cs % 4 == 0 and os.environ.get("GDN_FOLD_BF16", "1") != "0"
            tmp = torch.empty((N, d.out_h, d.out_w, cs), dtype=torch.bfloat16 if tmp16 else torch.float32, device=dev)
            if tmp16:
                d.out_bf16 = Act(tmp.data_ptr(), N, d.out_h, d.out_w, cs, 0)
            else:
                d.out_f32 = tmp.data_ptr()
            self.bwd.append(self._call(L.gdn_conv2d, d, "dgrad " + u.conv))
            self._dgrad_conv_end = len(self.bwd)
            f = FoldDesc()
            f.dpad_is_bf16 = int(tmp16)
            f.dpad = tmp.data_ptr()
            f.ctot, f.c_off = cs, 0
            f.n, f.h, f.w, f.c = N, h_s, w_s, cs
            f.pad, f.reflect, f.up, f.dilate = pad_phys, refl, up, dil
            f.dact = self.dact[s_name].data_ptr()
            f.accumulate = int(acc)
            self.bwd.append(self._call(L.gdn_fold_grad, f, "fold " + u.conv))
            self.launches_bwd += 2
            cu.keep = getattr(cu, "keep", []) + [tmp]
        cu.keep = getattr(cu, "keep", []) + [wdg]

    def _fuse_bn_backward_stats(self, d, u, s_name):
        """d: a single-launch input-gradient descriptor of unit u that writes d(s_name) directly.  If it is the LAST
        contribution to that gradient, let its epilogue reduce the BatchNorm-backward sums of the unit that produced
        s_name (and apply that unit's ReLU mask); the gradient is then kept as bf16 only unless the producer's identity
        branch still reads the fp32 value."""
        if not (self.train and self.do_bwd and getattr(self, "_fuse_bnbwd", False)):
            return
        p = self.producer.get(s_name)
        if p is None or p.bn is None or p.tanh or self._last_contrib.get(s_name) is not u:
            return
        if (p.relu and p.resid) or p.cout % 32 or d.cout != p.cout or d.dst_sy != 1 or d.dst_sx != 1:
            return
        # launches with a short reduction (1x1 convolutions on the concatenation, the im2col'd head) are bound by their
        # epilogue already: the extra epilogue work doubled their time on the B200 (profiles/r02c_profile_ops.log:
        # 0.113 -> 0.248 ms, more than the reduce pass it replaces) -- those keep the separate reduction
        if d.kh * d.kw * d.src0.c < int(os.environ.get("GDN_FUSE_BNBWD_MINK", "2048")):
            return
        cp = self.cu[p.conv]
        bn_i = self._bn_idx[p.conv]
        d.bwd_raw = cp.raw.data_ptr()
        d.bwd_coef = cp.coef4.data_ptr()
        d.bwd_relu = int(p.relu)
        d.stat_sum = self.bn_sums_all[bn_i, 0].data_ptr()
        d.stat_sqsum = self.bn_sums_all[bn_i, 1].data_ptr()
        self._fused.add(s_name)
        if not p.resid:
            # nothing but the producer's BatchNorm backward reads this gradient: bf16, ReLU mask already applied
            c, h, w = self.shape[s_name]
            self.gm[s_name] = torch.empty((self.N, h, w, c), dtype=torch.bfloat16, device=self.dev)
            d.out_f32 = None
            d.out_bf16 = Act(self.gm[s_name].data_ptr(), self.N, h, w, c, 0)

    def _build_dgrad_stride2(self, u, cu, s_name, c_off, cs, acc):
        """input gradient of a stride-2 convolution as FOUR stride-1 convolutions over dy (sub-pixel decomposition):
        the forward reads in[2*o + t + offb], so buffer positions Y with (Y - offb) = 2*i + a only ever meet the taps
        t = a + 2*m.  Per parity class (ay, ax):  dX[2*i + a + offb] = sum_m dy[i - m] * w[a + 2*m], a ceil((k-a)/2)-tap
        correlation with the flipped sub-kernel, written with destination stride 2.  No zero-dilated copy of dy and
        a quarter of the MMAs of the dilated formulation."""
        L, N, dev, P = self.L, self.N, self.dev, self.P
        k, kk = cu.k, cu.k * cu.k
        wt = P[u.conv + ".weight"]
        up, pad_phys, refl, dil = self._variant(u)
        assert not up and not dil
        c_s, h_s, w_s = self.shape[s_name]
        offb = cu.off + pad_phys
        Hd, Wd = h_s + 2 * pad_phys, w_s + 2 * pad_phys
        direct = not refl
        tmp16 = (not direct) and cs % 4 == 0 and os.environ.get("GDN_FOLD_BF16", "1") != "0"
        if direct:
            tgt = self.dact[s_name]
        else:
            tgt = torch.empty((N, Hd, Wd, cs), dtype=torch.bfloat16 if tmp16 else torch.float32, device=dev)
        keep = []
        for ay in (0, 1):
            for ax in (0, 1):
                ky, kx = (k - ay + 1) // 2, (k - ax + 1) // 2
                if ky <= 0 or kx <= 0:
                    continue
                wdg = torch.empty((ky * kx, cs, u.cout), dtype=torch.bfloat16, device=dev)
                # packed[(r,s)][ci][co] = w[co][c_off + ci][ay + 2*(ky-1-r)][ax + 2*(kx-1-s)]
                pd = PackDesc(ky, kx, cs, u.cout, cs, u.cout, kk, u.cin * kk, 2 * k, 2, 1, 0)
                self.pack_ops_bwd.append(self._pack_call(pd, wt, None, wdg, "pack-dgrad-s2 " + u.conv,
                                                         c_off * kk + ay * k + ax))
                iy = -((ay + offb) // 2)           # ceil((-a - offb) / 2): first i with a destination >= 0
                ix = -((ax + offb) // 2)
                oy, ox = 2 * iy + ay + offb, 2 * ix + ax + offb
                d = ConvDesc()
                d.src0 = Act(cu.dy.data_ptr(), N, cu.ho, cu.wo, u.cout, 0)
                d.weights = wdg.data_ptr()
                d.kh, d.kw = ky, kx
                d.stride = 1
                d.off_y, d.off_x = iy - (ky - 1), ix - (kx - 1)
                d.out_h, d.out_w = (Hd - oy + 1) // 2, (Wd - ox + 1) // 2
                d.cout = d.cout_pad = cs
                d.algo = 0
                d.dst_h, d.dst_w = Hd, Wd
                d.dst_sy = d.dst_sx = 2
                d.dst_oy, d.dst_ox = oy, ox
                if tmp16:
                    d.out_bf16 = Act(tgt.data_ptr(), N, Hd, Wd, cs, 0)
                else:
                    d.out_f32 = tgt.data_ptr()
                d.resid = tgt.data_ptr() if (direct and acc) else None
                if d.out_h > 0 and d.out_w > 0:
                    self.bwd.append(self._call(L.gdn_conv2d, d, "dgrad " + u.conv))
                    self._dgrad_conv_end = len(self.bwd)
                    self.launches_bwd += 1
                keep.append(wdg)
        if not direct:
            f = FoldDesc()
            f.dpad = tgt.data_ptr()
            f.dpad_is_bf16 = int(tmp16)
            f.ctot, f.c_off = cs, 0
            f.n, f.h, f.w, f.c = N, h_s, w_s, cs
            f.pad, f.reflect, f.up, f.dilate = pad_phys, refl, 0, 0
            f.dact = self.dact[s_name].data_ptr()
            f.accumulate = int(acc)
            self.bwd.append(self._call(L.gdn_fold_grad, f, "fold " + u.conv))
            self.launches_bwd += 1
            keep.append(tgt)
        cu.keep = getattr(cu, "keep", []) + keep

    def _build_head_backward(self, u, cu, have):
        """64 -> 1 head (k9, zero pad 4, tanh): the caller supplies dL/d(pre-tanh) as fp32 (N, H, W); it is im2col'd
        (81 taps -> 128 columns) so that both gradients run as 1x1 problems on the tensor cores."""
        L, N, dev, P = self.L, self.N, self.dev, self.P
        k, kk = cu.k, cu.k * cu.k
        ho, wo = cu.ho, cu.wo
        src = u.srcs[0]
        wt = P[u.conv + ".weight"]
        self.dpre = torch.zeros((N, ho, wo), dtype=torch.float32, device=dev)
        kp = _round_up(kk, 64)
        dcol = torch.empty((N, ho, wo, kp), dtype=torch.bfloat16, device=dev)
        cu.dcol = dcol

        def im2col(s):
            rc = L.gdn_im2col(C.c_void_p(self.dpre.data_ptr()), C.c_void_p(dcol.data_ptr()), N, 1, ho, wo, k, k, u.pad, 0, kp, s)
            if rc:
                _lib.check(rc, "im2col(dpre)")
        self.bwd.append(im2col)
        # input gradient: dX[q][c] = sum_t' dcol[q][t'] * Wd[c][t'],  Wd[c][t'] = w[c][flip(t')] (conv) / w[c][t'] (convT)
        wdg = torch.empty((1, u.cin, kp), dtype=torch.bfloat16, device=dev)
        pd = PackDesc(k, k, u.cin, kp, u.cin, kp, kk, 0, k, 1, 0 if u.transposed else 1, 1)
        self.pack_ops_bwd.append(self._pack_call(pd, wt, None, wdg, "pack-dgrad head"))
        d = ConvDesc()
        d.src0 = Act(dcol.data_ptr(), N, ho, wo, kp, 0)
        d.weights = wdg.data_ptr()
        d.kh = d.kw = 1
        d.stride = 1
        d.out_h, d.out_w = ho, wo
        d.cout = d.cout_pad = u.cin
        d.dst_h, d.dst_w = ho, wo
        d.dst_sy = d.dst_sx = 1
        acc = src in have
        have.add(src)
        tgt = self.dact[src]
        d.out_f32 = tgt.data_ptr()
        d.resid = tgt.data_ptr() if acc else None
        self._fuse_bn_backward_stats(d, u, src)
        self.bwd.append(self._call(L.gdn_conv2d, d, "dgrad head"))
        # weight gradient: dw[0][c][t'] = sum_q x[q][c] * dcol[q][t']
        v = self._variant(u)
        wd = WgradDesc()
        wd.x0 = self._act_struct(src, v)
        wd.dy = Act(dcol.data_ptr(), N, ho, wo, kp, 0)
        dw = self.dw_scratch[: u.cin * kp]
        wd.dw = dw.data_ptr()
        wd.kh = wd.kw = 1
        wd.stride = 1
        wd.out_h, wd.out_w, wd.cout_pad = ho, wo, kp
        used = self._det_slabs(wd)
        if used is None:
            self.bwd.append(lambda s, dw=dw: dw.zero_())
        self.bwd.append(self._call(L.gdn_conv2d_wgrad, wd, "wgrad head"))
        gbuf = self.grad[u.conv + ".weight"]
        # dw is [ci = c][co = t'] ; parameter is w[0][c][r][s] (conv, taps flipped) or w[c][0][r][s] (convT)
        if u.transposed:
            up_ = PackDesc(1, 1, kk, u.cin, kp, u.cin, 1, kk, 0, 0, 0, 0)
            goff = 0
        else:
            up_ = PackDesc(1, 1, kk, u.cin, kp, u.cin, -1, kk, 0, 0, 0, 0)
            goff = kk - 1

        def unpack(s):
            if used is None:
                rc = L.gdn_unpack_wgrad(C.byref(up_), C.c_void_p(dw.data_ptr()), C.c_void_p(gbuf.data_ptr() + 4 * goff), 1, s)
            else:
                rc = L.gdn_unpack_wgrad_slabs(C.byref(up_), C.c_void_p(self.dw_slabs.data_ptr()), used.value,
                                              C.c_int64(dw.numel()), C.c_void_p(gbuf.data_ptr() + 4 * goff), 1, s)
            if rc:
                _lib.check(rc, "unpack head")
        self.bwd.append(unpack)
        self.grad_ready_op[u.conv + ".weight"] = len(self.bwd) - 1
        self.launches_bwd += 5
        cu.keep = [wdg]

    def _attach_bn_finalize(self):
        """fuse every BatchNorm finalisation into the tail of its convolution (last CTA), see include/gdn_b200.h"""
        if not self.train or os.environ.get("GDN_FUSE_BNFIN", "1") == "0":
            return
        for cu in self.cu.values():
            fin = getattr(cu, "fin", None)
            d = getattr(cu, "conv_desc", None)
            if fin is None or d is None or ((d.algo >> 25) & 7) >= 2:
                continue
            slot, g_, b_, rm, rv, cnt = fin
            d.fin_counter = slot.data_ptr()
            d.fin_gamma, d.fin_beta = g_.data_ptr(), b_.data_ptr()
            d.fin_running_mean, d.fin_running_var = rm.data_ptr(), rv.data_ptr()
            d.fin_scale, d.fin_shift = cu.scale.data_ptr(), cu.shift.data_ptr()
            d.fin_mean, d.fin_rstd, d.fin_coef4 = cu.mean.data_ptr(), cu.rstd.data_ptr(), cu.coef4.data_ptr()
            d.fin_count, d.fin_eps, d.fin_momentum = cnt, BN_EPS, BN_MOMENTUM
            cu.fin_fused = True
            self.launches_fwd -= 1

    _DET_MAX_SLABS = 32

    def _det_slabs(self, wd):
        """GDN_DETERMINISTIC=1: point a weight-gradient descriptor at the engine's slab workspace (every split of the
        reduction stores its partial gradient to its own slab, the unpack sums them in order).  Returns the c_int32 the
        library writes the number of splits to at every call, or None in the default (atomic) mode."""
        if not self.det:
            return None
        if getattr(self, "dw_slabs", None) is None:
            self.dw_slabs = torch.empty(self._DET_MAX_SLABS * self.dw_scratch.numel(), dtype=torch.float32, device=self.dev)
        used = C.c_int32(0)
        wd.slabs = self.dw_slabs.data_ptr()
        wd.max_slabs = self._DET_MAX_SLABS
        wd.splits_used = C.pointer(used)
        self._keep_det = getattr(self, "_keep_det", []) + [used]
        return used

    def _attach_splitk_workspace(self):
        """one fp32 scratch buffer per engine for split-K launches (maps with fewer pixel tiles than SMs): every
        convolution of an engine runs on one stream at a time, so they can share it.  Sized for split 4 of the largest
        eligible launch; the library validates the rest (unsupported geometries return an error the autotuner skips)."""
        need = 0
        small = []
        for what, d in self._conv_descs:
            npix = d.src0.n * d.out_h * d.out_w
            if npix <= 128 * 148 and d.cout == d.cout_pad and d.cout >= 64 and d.dst_sy == 1 and d.dst_sx == 1:
                small.append(d)
                need = max(need, 4 * npix * d.cout * 4)
        self.splitk_ws = torch.empty(max(need, 16), dtype=torch.uint8, device=self.dev) if small else None
        for d in small:
            d.workspace = self.splitk_ws.data_ptr()
            d.workspace_bytes = self.splitk_ws.numel()

    def autotune(self, reps=3):
        """Time every staging variant of every implicit-GEMM launch once, on the device, and keep the fastest.  All
        variants accumulate in the same order, so the choice changes speed only (tile shape vs. wave quantisation
        on 148 SMs is what decides it: e.g. 16x52 maps of 512 channels run 1.7x faster tap-by-tap than halo-resident).
        Runs at engine construction, before any real data is in the buffers; GDN_AUTOTUNE=0 keeps the heuristics."""
        self.algo_choice = {}
        cache = {}
        # split-K is OPT-IN (GDN_SPLITK=1).  Measured on the B200 with graph-captured timing (profiles/r02f_sweep_conv.log):
        # the 512-channel 8x26 layers go from 47.0 to 43.1 us -- they are bound by the weight tiles every CTA pulls from L2
        # (2.3 MB per CTA at 128-byte granularity), not by the number of pixel tiles, so splitting the reduction moves
        # little; and a time-picked split changes the fp32 summation order, which would make two engines (two ranks, graph
        # vs eager, demo replay vs eager) differ in the last bits.  Without it every variant is bit-identical.
        if os.environ.get("GDN_SPLITK", "0") == "1":
            self._attach_splitk_workspace()
        for what, d in self._conv_descs:
            key = (d.src0.n, d.src0.h, d.src0.w, d.src0.c, d.src0.pad, d.src1.c if d.src1.ptr else 0, d.kh, d.kw, d.stride,
                   d.out_h, d.out_w, d.cout_pad, bool(d.out_f32), bool(d.out_bf16.ptr), bool(d.resid), bool(d.stat_sum),
                   d.dst_sy, bool(d.bwd_raw), bool(d.workspace), d.out_bf16.pad if d.out_bf16.ptr else 0)
            if key not in cache:
                cache[key] = autotune_conv(self.L, d, reps, what)
            d.algo = cache[key]
            self.algo_choice[what] = d.algo

    def profile(self, ops, reps=3, with_flops=False):
        """per-op device time (ms; each op captured in a CUDA graph and replayed) of a list of ops (self.fwd or self.bwd).
        with_flops: rows are (label, ms, flops or None, 128-pixel tiles or None)"""
        res = []
        for op in ops:
            ms = graph_time_ms(op, reps)       # graph-captured: eager timing of the small ops measures the host
            if with_flops:
                res.append((getattr(op, "label", "misc"), ms, getattr(op, "flops", None), getattr(op, "tiles", None)))
            else:
                res.append((getattr(op, "label", "misc"), ms))
        return res

    def backward(self, dpre=None, dout_nhwc=None):
        """dpre: dL/d(pre-tanh output), fp32 (N, H, W) -- or None if the loss kernel already wrote self.dpre.
        dout_nhwc: for graphs without a tanh head (stand-alone blocks), the gradient of the last tensor.
        Parameter gradients are ACCUMULATED into self.grad[...] (views of self.flat_grad)."""
        if dpre is not None:
            self.dpre.copy_(dpre.reshape(self.dpre.shape))
        if dout_nhwc is not None:
            self.dact[self.units[-1].out].copy_(dout_nhwc)
        self.run_backward()

    def run_backward(self, after_op=None):
        """enqueue the backward plan: main-chain ops on the current stream, the weight-gradient groups on the side
        stream (each group waits for the main stream's position at its place in the plan; the main stream waits for
        the side stream at the end).  after_op(i) is called after op i has been enqueued (gradient bucket launches)."""
        s = _lib.stream_ptr()
        side = self.side_stream
        if side is None:
            for i, op in enumerate(self.bwd):
                op(s)
                if after_op is not None:
                    after_op(i)
            return
        main = torch.cuda.current_stream(self.dev)
        s_side = C.c_void_p(side.cuda_stream)
        pack_ev = getattr(self, "_pack_ev", None)
        if pack_ev is not None:         # dgrad weight packs issued on the side stream during forward
            main.wait_event(pack_ev)
            self._pack_ev = None
        side.wait_stream(main)          # fork now: the side stream is part of the step (and of a graph capture) from here on
        prev_side = False
        tl = getattr(self, "timeline", None)
        for i, op in enumerate(self.bwd):
            if getattr(op, "side", 0):
                if not prev_side:
                    ev = torch.cuda.Event()
                    ev.record(main)
                    side.wait_event(ev)
                with torch.cuda.stream(side):
                    op(s_side)
                prev_side = True
            else:
                op(s)
                prev_side = False
            if tl is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(side if prev_side else main)
                tl.append((getattr(op, "label", "misc"), "bwd-side" if prev_side else "bwd-main", ev))
            if after_op is not None:
                after_op(i)
        main.wait_stream(side)


# ------------------------------------------------------------------------------------------- conv autotuning
# algo word of gdn_conv_desc: bits 0-7 mode (GDN_CONV_TAPBOX = 1, GDN_CONV_HALO = 2), bits 8-15 HALO sub-tiles J,
# bits 16-23 output-channel tile / 64 (0 = widest), bit 24 CTA pairs (tcgen05 cta_group::2), bit 28 second epilogue warp group
_PAIR = 1 << 24
_EW8 = 1 << 28
_ALGOS = (2 | (4 << 8), 2 | (2 << 8), 2 | (1 << 8), 1,
          2 | (4 << 8) | _PAIR, 2 | (2 << 8) | _PAIR, 2 | (1 << 8) | _PAIR, 1 | _PAIR)
_S2, _S4 = 2 << 25, 4 << 25     # split-K over the input-channel chunks (needs the engine's workspace)
_ALGOS_SPLIT = (1 | _PAIR | _S2, 1 | _PAIR | _S4, 1 | _S2, 1 | _S4, 1 | (2 << 16) | _PAIR | _S2, 1 | (2 << 16) | _S2,
                2 | (1 << 8) | _PAIR | _S2, 2 | (1 << 8) | _S2, 2 | (2 << 8) | (2 << 16) | _PAIR | _S2)
_ALGOS_NARROW = (2 | (2 << 8) | (2 << 16), 2 | (1 << 8) | (2 << 16), 1 | (2 << 16),   # 128-wide channel tiles
                 2 | (2 << 8) | (2 << 16) | _PAIR, 2 | (1 << 8) | (2 << 16) | _PAIR, 1 | (2 << 16) | _PAIR)


def autotune_conv(L, d, reps=3, what="conv"):
    """fastest algo word for one gdn_conv_desc (timed on the device with CUDA events); leaves d.algo set to it"""
    s = _lib.stream_ptr()
    fn = L.gdn_conv2d
    pairs_ok = os.environ.get("GDN_PAIRS", "1") != "0"
    best, best_ms = 0, None
    # wide layers on small maps leave SMs idle with 256-channel tiles: let 128-wide tiles compete
    small = d.cout_pad >= 256 and d.src0.n * d.out_h * d.out_w * (d.cout_pad // 256) < 128 * 148 * 2
    split = _ALGOS_SPLIT if (d.workspace and d.src0.n * d.out_h * d.out_w <= 128 * 148) else ()
    cands = _ALGOS + (_ALGOS_NARROW if small else ()) + split
    # launches with a short reduction are bound by their epilogue (one warp per scheduler: ~2 900 cycles per 32-column group
    # against K = Cin*kh*kw cycles of MMAs): let the variants with eight epilogue warps compete there
    # (measured, profiles/r02m_sweep_conv.log: 128->64 k1 172 -> 144 us, 64->128 k4 s2 166 -> 132 us, 512-channel k3 on
    # 16x52 67.6 -> 64.6 us, and still 151.8 -> 149.0 us on 256 -> 256 k5 with K = 6400 -- so every launch may try them)
    if d.cout_pad >= 64 and os.environ.get("GDN_EW8", "1") != "0":
        cands = cands + tuple(a | _EW8 for a in cands)
    for algo in cands:
        if (algo & _PAIR) and not pairs_ok:
            continue
        d.algo = algo
        if fn(C.byref(d), s) != 0:       # variant not applicable to this geometry
            continue
        if _SYNC_DEBUG:
            try:
                torch.cuda.synchronize()
            except Exception as e:
                raise RuntimeError("gdn_b200: kernel failure autotuning '%s' with algo 0x%x: %s" % (what, algo, e))
        ms = graph_time_ms(lambda sp: fn(C.byref(d), sp), reps)
        if best_ms is None or ms < best_ms:
            best, best_ms = algo, ms
    d.algo = best
    return best


def graph_time_ms(launch, reps=3, inner=6):
    """device time of ``inner`` back-to-back launches, captured in a CUDA graph and replayed ``reps`` times (ms per launch,
    best replay).  Eager timing of small kernels measures the HOST: one gdn_conv2d call costs ~40 us of Python / ctypes /
    tensor-map encoding, more than the kernel itself on the 8x26 and 16x52 maps -- the round-1 autotuner and the round-2
    variant sweeps (profiles/r02b_sweep_conv.log: every variant of the small layers at the same ~47 us) were blind there."""
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        sp = _lib.stream_ptr()
        for _ in range(inner):
            launch(sp)
    g.replay()                               # warm-up replay (graph upload)
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / inner
        best = ms if best is None else min(best, ms)
    del g
    return best
