"""Fused training steps: the per-iteration bodies of the reference's train loops on B200 kernels.

  RtoDTrainStep  = /root/reference/src/trainer.py:696-768  (train_AE_RtoD: model fwd, two frozen no-grad DtoD passes,
                   masked BerHu + latent MSE + edge-aware smoothness, zero_grad / backward / Adam.step)
  DtoDTrainStep  = /root/reference/src/trainer.py:427-468  (train_AE_DtoD: model fwd, BerHu + 3*Sobel loss, Adam)
  Optimizer      = optim.Adam(params, lr, [0.9, 0.999], eps=1e-8, weight_decay=5e-4), GDN_main.py:157,173
  Data parallel  = replaces nn.DataParallel (GDN_main.py:153-198): one process per GPU, parameters replicated,
                   per-shard BatchNorm statistics (what DataParallel replicas do), gradient all-reduce (AVG) over
                   NCCL in ~25 MB buckets launched while the rest of backward is still running, plus one 4-byte
                   all-reduce(MAX) so the BerHu threshold c = 0.2*max|diff| stays a GLOBAL-batch quantity as in the
                   reference (loss computed on the gathered batch, trainer.py:711-720).

Nothing here synchronises with the host; losses are returned as 0-dim device tensors.
"""
import os

import torch
import torch.distributed as dist

from . import _lib
from .module_runtime import get_engine, _engines, _after_train_forward
from .ops import LossKernels, FusedAdam


def _round_up(v, m):
    return (v + m - 1) // m * m


def flatten_parameters(module):
    """Re-home every parameter of ``module`` into ONE flat fp32 buffer (4-element aligned slots, named_parameters
    order) so that Adam and the gradient all-reduce are single contiguous operations.  Values are preserved;
    state_dict()/load_state_dict() keep working (they copy in place).  Returns the flat buffer."""
    named = list(module.named_parameters())
    total = sum(_round_up(p.numel(), 4) for _, p in named)
    dev = named[0][1].device
    flat = torch.zeros(total, dtype=torch.float32, device=dev)
    o = 0
    with torch.no_grad():
        for _, p in named:
            n = p.numel()
            view = flat[o:o + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
            o += _round_up(n, 4)
    module.__dict__.pop("_gdn_engines", None)   # engines hold raw pointers to the old storage
    return flat


class GradBuckets:
    """Contiguous ~bucket_bytes slices of the flat gradient buffer, ordered by when they become final during
    backward.  ``ready_after[i]`` is the index of the backward op after which bucket i may be all-reduced."""

    def __init__(self, slots, ready_op, bucket_bytes=25 << 20):
        """slots: [(name, offset, numel_padded)] in buffer order; ready_op: {name: op index after which final}"""
        self.buckets = []  # (start, end, ready_after)
        cur_s, cur_e, cur_r = None, None, -1
        for name, off, n in slots:
            if cur_s is None:
                cur_s, cur_e, cur_r = off, off + n, ready_op.get(name, -1)
            else:
                cur_e = off + n
                cur_r = max(cur_r, ready_op.get(name, -1))
            if (cur_e - cur_s) * 4 >= bucket_bytes:
                self.buckets.append((cur_s, cur_e, cur_r))
                cur_s = None
        if cur_s is not None:
            self.buckets.append((cur_s, cur_e, cur_r))
        self.buckets.sort(key=lambda b: b[2])


def allreduce_avg_(flat, buckets, group=None, async_streams=None):
    """in-place average of ``flat`` across the group, bucket by bucket (host logic shared by the NCCL path and the
    gloo CPU tests)"""
    world = dist.get_world_size(group)
    works = []
    for s, e, _ in buckets:
        works.append((dist.all_reduce(flat[s:e], op=dist.ReduceOp.SUM, group=group, async_op=True), s, e))
    for w, s, e in works:
        w.wait()
    flat.mul_(1.0 / world)
    return flat


class _StepBase:
    """shared machinery of the fused steps: flat parameters, bucketed all-reduce, fused Adam, CUDA-graph replay.

    CUDA graph: after ``graph_warmup`` eager steps the whole step (weight re-pack, forward, guidance passes, loss,
    backward, Adam) is captured ONCE and then replayed -- ~600 kernel launches become one cudaGraphLaunch, which
    removes the host from the critical path.  Step-dependent Adam scalars and the learning rate live in a 2-float
    device buffer refreshed before each replay.  With world > 1 the bucketed NCCL all-reduces (issued on the side
    stream, forked from and joined back into the capturing stream by events) are part of the same graph.
    Set GDN_GRAPH=0 to run eagerly."""
    graph_warmup = 2

    def __init__(self, model, lr, betas, eps, weight_decay, group, bucket_mb):
        self.model = model
        self.dev = next(model.parameters()).device
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        if self.world > 1:
            # replicate rank 0's parameters and buffers once (DataParallel re-broadcasts them every forward)
            with torch.no_grad():
                for t in list(model.parameters()) + list(model.buffers()):
                    dist.broadcast(t.data, src=0, group=group)
        self.flat_params = flatten_parameters(model)
        self.graph = model.gdn_graph()
        model.__dict__["_gdn_graph"] = self.graph
        self.kern = LossKernels(self.dev)
        self.lr = lr
        self.hyper = (betas, eps, weight_decay)
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        self.eng = None
        self.opt = None
        self.comm_stream = torch.cuda.Stream(device=self.dev) if self.world > 1 else None
        # Stream priorities: the step's critical chain (forward, dgrad / BatchNorm-backward chain, Adam) runs on a
        # HIGH priority stream, the weight-gradient chain on a middle one (engine.side_stream) and the frozen DtoD
        # guidance passes on a low one: pending CTAs of the critical chain are always placed first, the rest fills
        # whatever SMs it leaves idle (its HBM-bound passes, tails of its persistent kernels).
        self.main_stream = None
        if os.environ.get("GDN_SIDE", "1") != "0":
            self.main_stream = torch.cuda.Stream(device=self.dev, priority=-2)
        self.step_count = 0
        self.launches_per_step = 0
        env = os.environ.get("GDN_GRAPH")
        self.use_graph = (env != "0")                # also with world > 1: NCCL collectives are captured into the graph
        self._graph = None
        self._static_in = None
        self._static_out = None

    def _ensure(self, x):
        if self.eng is not None and self.eng.N == x.shape[0] and self.eng.H == x.shape[2] and self.eng.W == x.shape[3]:
            return
        self.model.train()
        _engines(self.model)
        self.eng = get_engine(self.model, self.graph, x, train=True, backward=True, want=())
        eng = self.eng
        eng.async_bwd_pack = True
        eng.async_fwd_pack = True
        named = list(self.model.named_parameters())
        # the engine's flat gradient uses the same slot layout as flatten_parameters()
        assert eng.flat_grad.numel() == self.flat_params.numel(), (eng.flat_grad.numel(), self.flat_params.numel())
        if self.opt is None:
            betas, eps, wd = self.hyper
            self.opt = FusedAdam([p for _, p in named], lr=self.lr, betas=betas, eps=eps, weight_decay=wd,
                                 flat=(self.flat_params, eng.flat_grad))
        else:
            self.opt.flat = (self.flat_params, eng.flat_grad)
        pend = self.__dict__.pop("_pending_opt_state", None)
        if pend is not None:
            self._apply_opt_state(pend)
        # bucket plan: a parameter's gradient is final after the last backward op that mentions it
        slots, o = [], 0
        for n, p in named:
            slots.append((n, o, _round_up(p.numel(), 4)))
            o += _round_up(p.numel(), 4)
        ready = getattr(eng, "grad_ready_op", {})
        self.buckets = GradBuckets(slots, ready, self.bucket_bytes).buckets

    def set_lr(self, lr):
        """the reference mutates param_groups[...]['lr'] (trainer.py:784-792); same thing here"""
        self.lr = lr
        if self.opt is not None:
            self.opt.set_lr(lr)

    def _backward_and_reduce(self):
        eng = self.eng
        s = _lib.stream_ptr()
        eng.flat_grad.zero_()
        if self.world == 1:
            eng.run_backward()
            return
        cur = torch.cuda.current_stream(self.dev)
        bi = 0
        nb = len(self.buckets)
        if getattr(self, "debug_keep_local_grad", False):
            # test hook (tools/check_ddp.py): finish backward first and keep this shard's own gradient
            eng.run_backward()
            self.local_grad = eng.flat_grad.clone()
            for bi in range(nb):
                self._launch_bucket(bi, cur)
            cur.wait_stream(self.comm_stream)
            return
        state = [0]

        def after_op(i):
            while state[0] < nb and self.buckets[state[0]][2] <= i:
                self._launch_bucket(state[0], cur)
                state[0] += 1
        eng.run_backward(after_op)
        while state[0] < nb:
            self._launch_bucket(state[0], cur)
            state[0] += 1
        cur.wait_stream(self.comm_stream)

    def _launch_bucket(self, bi, cur):
        s0, e0, _ = self.buckets[bi]
        ev = torch.cuda.Event()
        ev.record(cur)
        self.comm_stream.wait_event(ev)
        side = getattr(self.eng, "side_stream", None)
        if side is not None:                 # weight gradients are produced on the engine's side stream
            ev2 = torch.cuda.Event()
            ev2.record(side)
            self.comm_stream.wait_event(ev2)
        with torch.cuda.stream(self.comm_stream):
            dist.all_reduce(self.eng.flat_grad[s0:e0], op=dist.ReduceOp.SUM, group=self.group)

    def _run(self, inputs):
        """eager for the first steps, then capture-and-replay"""
        shapes = tuple(None if t is None else tuple(t.shape) for t in inputs)
        if not self.use_graph:
            self.step_count += 1
            return self._eager_on_main(inputs)
        if self._graph is not None and shapes != self._graph_shapes:
            self._graph = None          # new batch shape: fall back to eager warm-up and re-capture
            self._eager_done = 0
        if self._graph is None and getattr(self, "_eager_done", 0) < self.graph_warmup:
            self._eager_done = getattr(self, "_eager_done", 0) + 1
            self.step_count += 1
            return self._eager_on_main(inputs)
        self.step_count += 1
        if self._graph is None:
            self._static_in = [None if t is None else t.clone() for t in inputs]
            self._graph_shapes = shapes
            torch.cuda.synchronize(self.dev)
            if self.world > 1:
                dist.barrier(group=self.group)       # every rank's eager collectives have retired before capture
                torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            # thread_local: the NCCL watchdog thread may touch the CUDA API while this thread captures
            kw = {"stream": self.main_stream} if self.main_stream is not None else {}
            with torch.cuda.graph(g, capture_error_mode="thread_local" if self.world > 1 else "global", **kw):
                self._static_out = self._eager(*self._static_in)
            self._graph = g
        for st, t in zip(self._static_in, inputs):
            if st is not None and st.data_ptr() != t.data_ptr():
                st.copy_(t, non_blocking=True)
        self._graph.replay()
        # the replayed Adam / BatchNorm-finalize kernels rewrote parameters and running statistics through raw pointers
        # (no tensor _version bump, none of the Python bookkeeping of _optim_step / _after_train_forward runs):
        # every OTHER engine of this module (cached eval-mode engines, DeviceValidator) must re-pack / re-fold
        self.model.__dict__["_gdn_epoch"] = self.model.__dict__.get("_gdn_epoch", 0) + 1
        return self._static_out

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self):
        """Everything needed to resume training exactly: the model's own state_dict (keys identical to the
        reference's checkpoints, trainer.py:518,843 -- it can be loaded by the reference as is), the Adam moments
        (flat, named_parameters order), the step count and the current learning rate.  The reference saves weights
        only (no optimizer state, no resume flag: SURVEY.md section 5)."""
        sd = {"model": {k: v.detach().clone() for k, v in self.model.state_dict().items()},
              "step": self._steps_taken(), "lr": float(self.lr),
              "hyper": {"betas": tuple(self.hyper[0]), "eps": self.hyper[1], "weight_decay": self.hyper[2]}}
        if self.opt is not None and self.opt._flat_state is not None:
            sd["adam_m"], sd["adam_v"] = (t.detach().clone() for t in self.opt._flat_state)
        return sd

    def _steps_taken(self):
        """optimizer steps so far (under graph replay the authoritative counter is the device-side one)"""
        if self.opt is None:
            return 0
        if self.opt.dyn is not None:
            return int(round(float(self.opt.dyn[1])))
        return int(self.opt._step)

    def load_state_dict(self, sd):
        """inverse of state_dict(); parameters are copied IN PLACE (the flat buffer and every raw pointer the
        engines hold stay valid), the CUDA graph keeps replaying with the restored step count / learning rate"""
        hy = sd.get("hyper")
        if hy is not None:
            saved = (tuple(float(b) for b in hy["betas"]), float(hy["eps"]), float(hy["weight_decay"]))
            mine = (tuple(float(b) for b in self.hyper[0]), float(self.hyper[1]), float(self.hyper[2]))
            if saved != mine:
                if self._graph is not None:
                    raise RuntimeError("gdn_b200: checkpoint Adam hyper-parameters %s differ from this step's %s and the step "
                                       "is already captured in a CUDA graph; construct the step with the saved values"
                                       % (saved, mine))
                self.hyper = (saved[0], saved[1], saved[2])     # resume exactly: the checkpoint's betas / eps / decay win
                if self.opt is not None:
                    self.opt.set_hyper(*self.hyper)
        self.model.load_state_dict(sd["model"])
        self.model.__dict__["_gdn_epoch"] = self.model.__dict__.get("_gdn_epoch", 0) + 1   # engines re-pack / re-fold
        if self.eng is not None:
            self.eng._wversion = None
        self.lr = float(sd["lr"])
        if self.opt is None:
            self._pending_opt_state = sd       # applied when the optimizer is created by the first step
            return
        self._apply_opt_state(sd)

    def _apply_opt_state(self, sd):
        opt = self.opt
        opt._step = int(sd["step"])
        opt.set_lr(float(sd["lr"]))
        if "adam_m" in sd:
            if opt._flat_state is None:
                opt._flat_state = (torch.zeros_like(self.flat_params), torch.zeros_like(self.flat_params))
            opt._flat_state[0].copy_(sd["adam_m"])
            opt._flat_state[1].copy_(sd["adam_v"])
        if opt.dyn is not None:
            opt.dyn[1:2].fill_(float(opt._step))

    def save_checkpoint(self, path):
        torch.save(self.state_dict(), path)

    def load_checkpoint(self, path):
        self.load_state_dict(torch.load(path, map_location=self.dev))

    def close(self):
        """Release everything that pins the process group: the captured CUDA graph (it holds the NCCL kernels of the
        bucketed all-reduces -- a communicator cannot be torn down while a live graph still references its kernels), the
        static input / output buffers and the engines.  Call before dist.destroy_process_group(); the step object can be
        used again afterwards (it re-captures)."""
        import gc
        torch.cuda.synchronize(self.dev)
        self._graph = None
        self._static_in = self._static_out = None
        self._eager_done = 0
        gc.collect()
        torch.cuda.synchronize(self.dev)

    def _eager_on_main(self, inputs):
        """run the step on the high-priority stream, ordered after / before the caller's current stream"""
        ms = self.main_stream
        if ms is None:
            return self._eager(*inputs)
        cur = torch.cuda.current_stream(self.dev)
        ms.wait_stream(cur)
        with torch.cuda.stream(ms):
            out = self._eager(*inputs)
        cur.wait_stream(ms)
        return out

    def _optim_step(self):
        self.opt.grad_scale = 1.0 / self.world     # SUM all-reduce -> average, folded into the Adam kernel
        # learning rate and step count always live in device memory (what a captured graph needs), also when the step runs
        # eagerly: both launch modes then evaluate the bias corrections with the same arithmetic, and in deterministic mode
        # a CUDA-graph run is bit-identical to an eager one (tools/check_deterministic.py)
        self.opt.enable_device_step()
        self.opt.step()
        self.eng._wversion = None                    # parameters changed through raw pointers: re-pack next forward
        self.model.__dict__["_gdn_epoch"] = self.model.__dict__.get("_gdn_epoch", 0) + 1   # ... in every other engine too


class DtoDTrainStep(_StepBase):
    """one iteration of train_AE_DtoD (trainer.py:411-468)"""

    def __init__(self, model, lr=2e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=5e-4, group=None, bucket_mb=25):
        super().__init__(model, lr, betas, eps, weight_decay, group, bucket_mb)

    def step(self, depths, sparse):
        """depths: (N,1,H,W) dense gt in [-1,1]; sparse: (N,1|3,H,W) sparse gt (invalid = -1) or None."""
        return self._run((depths, sparse))

    def _eager(self, depths, sparse):
        self._ensure(depths)
        eng = self.eng
        eng.forward(depths)
        _after_train_forward(self.model)
        out = eng.depth()
        self.kern.absdiff_max(out, depths)
        if self.world > 1:
            dist.all_reduce(self.kern.maxabs, op=dist.ReduceOp.MAX, group=self.group)
        self.kern.loss(1, out, depths, sparse, None, dpre=eng.dpre)
        self._backward_and_reduce()
        self._optim_step()
        return self.kern.assemble(1, float(out.numel()))


class RtoDTrainStep(_StepBase):
    """one iteration of train_AE_RtoD (trainer.py:670-768) with a frozen, eval-mode DtoD guidance network"""

    def __init__(self, model, dtod_model, lr=2e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=5e-4, group=None,
                 bucket_mb=25, guidance=True, guidance_grad=False):
        """guidance_grad=False (default) reproduces the published code: the latent loss is a reported constant
        (both DtoD passes run under no_grad, trainer.py:699-703).  guidance_grad=True is the paper-faithful opt-in
        (SURVEY.md 8f row 3): the latent loss back-propagates through the frozen DtoD encoder into the RtoD output
        (input gradients of the DtoD encoder only -- its weights stay frozen) and on through the RtoD network."""
        super().__init__(model, lr, betas, eps, weight_decay, group, bucket_mb)
        self.dtod = dtod_model
        self.guidance = guidance and dtod_model is not None
        self.guidance_grad = bool(guidance_grad) and self.guidance
        self.aux_stream = None
        if self.guidance:
            self.dtod.eval()
            self.dgraph = self.dtod.gdn_graph()
            self.deng = [None, None]
            if os.environ.get("GDN_SIDE", "1") != "0":
                # the target-feature pass depends only on the ground truth: it runs on its own stream, its
                # tensor-core kernels fill the SMs while the RtoD forward is in its HBM-bound BatchNorm passes
                self.aux_stream = torch.cuda.Stream(device=self.dev)

    def _dtod_features(self, slot, x):
        """encoder + bottleneck of the frozen DtoD net only: (x1, x2, x4, x6) is all the loss reads (trainer.py:700,703);
        two engine instances so the target features and the prediction features coexist"""
        e = self.deng[slot]
        if e is None or e.N != x.shape[0] or e.H != x.shape[2] or e.W != x.shape[3]:
            from .engine import Engine
            from .module_runtime import _params
            names = self.dgraph.encoder_outputs
            gg = self.guidance_grad and slot == 1     # the prediction pass carries the guidance gradient
            e = Engine(self.dgraph, _params(self.dtod), x.shape[0], x.shape[2], x.shape[3], train=False, backward=gg,
                       want=names, stop_after=names[-1], device=x.device, input_grad=gg, grad_seeds=names if gg else ())
            self.deng[slot] = e
        e.forward(x)
        return [e.value(n) for n in self.dgraph.encoder_outputs]

    def step(self, rgb, depths, sparse):
        """rgb (N,3,H,W), depths (N,1,H,W), sparse (N,1|3,H,W) or None; all fp32 in [-1,1] on the GPU."""
        return self._run((rgb, depths, sparse))

    def _eager(self, rgb, depths, sparse):
        self._ensure(rgb)
        eng = self.eng
        ft_tar = None
        main = torch.cuda.current_stream(self.dev)
        eng.refresh_if_stale()      # forward weight packs first: they are at the head of the critical chain
        if self.guidance and self.aux_stream is not None:
            self.aux_stream.wait_stream(main)
            with torch.cuda.stream(self.aux_stream), torch.no_grad():
                ft_tar = self._dtod_features(0, depths)
        eng.forward(rgb)
        _after_train_forward(self.model)
        out = eng.depth()
        feat_numels = None
        self.kern.absdiff_max(out, depths)
        if self.world > 1:
            dist.all_reduce(self.kern.maxabs, op=dist.ReduceOp.MAX, group=self.group)
        self.kern.loss(0, out, depths, sparse, rgb, dpre=eng.dpre)
        if self.guidance_grad:
            # paper-faithful opt-in: d latent / d out through the frozen DtoD encoder, chained onto dL/d(pre-tanh)
            # before the RtoD backward starts (so this pass IS on the critical chain, on the main stream)
            if ft_tar is None:
                ft_tar = self._dtod_features(0, depths)
            else:
                main.wait_stream(self.aux_stream)
            ft = self._dtod_features(1, out)
            self.kern.latent(ft, ft_tar)
            feat_numels = [float(t.numel()) for t in ft]
            e = self.deng[1]
            e.refresh_if_stale()
            self.kern.latent_grad(ft, ft_tar, [e.dact[n] for n in self.dgraph.encoder_outputs])
            e.run_backward()
            self.kern.tanh_chain_add(e.dact["in"], out, eng.dpre)
        elif self.guidance and ft_tar is None:
            with torch.no_grad():
                ft_tar = self._dtod_features(0, depths)
                ft = self._dtod_features(1, out)
            self.kern.latent(ft, ft_tar)
            feat_numels = [float(t.numel()) for t in ft]
        elif self.guidance:
            # the prediction-feature pass and the four feature MSEs feed only the REPORTED latent loss (no gradient:
            # trainer.py:699-703 runs them under no_grad), so they stay on the low-priority stream, under backward
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(self.aux_stream), torch.no_grad():
                self.aux_stream.wait_event(ev)
                ft = self._dtod_features(1, out)
                self.kern.latent(ft, ft_tar)
            feat_numels = [float(t.numel()) for t in ft]
        self._backward_and_reduce()
        self._optim_step()
        if self.aux_stream is not None:
            main.wait_stream(self.aux_stream)
        return self.kern.assemble(0, float(out.numel()), feat_numels)


def init_distributed_from_env():
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*); returns (rank, world, device)"""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
    else:
        dev = torch.device("cpu")
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", rank=rank, world_size=world)
    return rank, world, dev


def shutdown_distributed(steppers=(), timeout_s=30.0):
    """Orderly end of a multi-rank run: drop the CUDA graphs that hold captured NCCL kernels (``close()``), barrier,
    destroy the process group.  Round 1 dodged a teardown dead-lock behind live graphs with os._exit(0); the dead-lock
    was the graph outliving the communicator.  A watchdog keeps the old escape hatch: if teardown has not finished after
    ``timeout_s`` the process still exits with status 0 (and says so on stderr) instead of hanging a benchmark.
    Returns True when the group was destroyed cleanly."""
    import sys
    import threading
    if not (dist.is_available() and dist.is_initialized()):
        return True
    done = threading.Event()

    def work():
        for st in steppers:
            if st is not None:
                st.close()
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        done.set()
    t = threading.Thread(target=work, daemon=True)
    t.start()
    t.join(timeout_s)
    if not done.is_set():
        sys.stderr.write("gdn_b200: process-group teardown did not finish in %.0f s; leaving with os._exit(0)\n" % timeout_s)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return True
