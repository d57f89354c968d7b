"""Glue between the nn.Module facade (AE_model_unet.py) and the engine: engine caching, autograd integration.

Semantics kept from the reference call sites: ``model(x, istrain=False)`` -> depth (N,1,H,W) fp32;
``istrain=True`` -> the 8-tuple; ``model.train()/.eval()`` select batch-statistics / running-statistics BatchNorm;
parameters receive ``.grad`` through ``loss.backward()`` like any other module; BN running stats and
``num_batches_tracked`` are updated in train mode.
"""
import torch

from .engine import Engine
from .graph import Graph, Unit


def _require_cuda(x):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError("gdn_b200: this is a CUDA-only (sm_100a) implementation; got a CPU tensor and there is no "
                           "PyTorch/CPU fallback")
    if x.dtype != torch.float32:
        raise RuntimeError("gdn_b200: inputs must be fp32 (the reference feeds fp32 tensors in [-1, 1])")


def _check_norm(module):
    """norm != 'Batch' builds nn.InstanceNorm2d(affine=True, track_running_stats=True) in the ConvBlocks / ConvTBlocks
    (reference AE_model_unet.py:70-76, 88-93).  In eval mode such a layer normalises with its RUNNING statistics
    (torch: use_input_stats = training or not track_running_stats), i.e. exactly like an eval-mode BatchNorm with the
    same state_dict keys -- it folds into the weights like one, so inference works.  In train mode it needs
    per-sample statistics, which no live configuration uses (option.py:43 defaults to Batch) and no kernel here
    computes: that is an error, not a silent substitution."""
    has_in = module.__dict__.get("_gdn_has_instance_norm")
    if has_in is None:
        has_in = any(isinstance(m, torch.nn.InstanceNorm2d) for m in module.modules())
        module.__dict__["_gdn_has_instance_norm"] = has_in
    if has_in and module.training:
        raise NotImplementedError("gdn_b200: norm='Instance' is implemented for eval-mode inference only (running "
                                  "statistics, like the reference in .eval()); train mode needs per-sample statistics "
                                  "-- the reference default is norm='Batch' (option.py:43) and no live configuration "
                                  "trains the InstanceNorm branch")


def _refuse_replica(module):
    """nn.DataParallel over several GPUs calls forward on replicas made by _replicate_for_data_parallel: their
    _parameters are empty (state_dict() holds buffers only), the parameters are non-leaf broadcast views, and the
    shallow __dict__ copy shares this runtime's engine cache between devices.  That path is what the fused,
    one-process-per-GPU steps replace (GDN_main.py:153-198 -> trainer.py here); refuse it loudly."""
    if getattr(module, "_is_replica", False):
        raise RuntimeError("gdn_b200: this module is an nn.DataParallel replica (several GPUs visible to one process). "
                           "Multi-GPU runs use one process per GPU: launch with torchrun and use "
                           "gdn_pytorch_b200.trainer.RtoDTrainStep / DtoDTrainStep (INTEGRATION.md section 3), or wrap as "
                           "nn.DataParallel(model, device_ids=[local_rank]) to keep the reference's checkpoint keys on one "
                           "device per process")


def _engines(module):
    _refuse_replica(module)
    d = module.__dict__.get("_gdn_engines")
    if d is None:
        d = {}
        module.__dict__["_gdn_engines"] = d
        module.__dict__["_gdn_epoch"] = 0
        _check_norm(module)
    return d


def _params(module):
    return dict(module.state_dict(keep_vars=True))


def _nbt(module):
    return [b for n, b in module.named_buffers() if n.endswith("num_batches_tracked")]


def get_engine(module, graph, x, train, backward, want=(), stop_after=None):
    N, _, H, W = x.shape
    key = (N, H, W, bool(train), bool(backward), tuple(want), stop_after, x.device.index)
    cache = _engines(module)
    eng = cache.get(key)
    if eng is None:
        with torch.cuda.device(x.device):
            eng = Engine(graph, _params(module), N, H, W, train=train, backward=backward, want=want,
                         stop_after=stop_after, device=x.device)
        eng._seen_epoch = -1
        cache[key] = eng
    if eng._seen_epoch != module.__dict__["_gdn_epoch"]:
        eng._wversion = None           # BN running statistics were rewritten by a train-mode forward
        eng._seen_epoch = module.__dict__["_gdn_epoch"]
    return eng


def _after_train_forward(module):
    module.__dict__["_gdn_epoch"] += 1
    nbt = _nbt(module)
    if nbt:
        torch._foreach_add_(nbt, 1)


class _NetFn(torch.autograd.Function):
    """whole-network forward/backward as ONE autograd node (the engine keeps its own activations)"""

    @staticmethod
    def forward(ctx, x, module, graph, istrain, names, *params):
        want = graph.outputs if istrain else ()
        eng = get_engine(module, graph, x, train=True, backward=True, want=want)
        eng.forward(x)
        eng._fwd_id = getattr(eng, "_fwd_id", 0) + 1
        _after_train_forward(module)
        ctx.eng, ctx.fwd_id, ctx.names = eng, eng._fwd_id, names
        depth = eng.depth().clone()
        ctx.save_for_backward(depth)
        ctx.set_materialize_grads(False)
        if istrain:
            # the feature maps stay differentiable outputs so that a loss on them is REFUSED in backward() instead of
            # silently contributing nothing (the reference would back-propagate it)
            feats = [eng.value_nchw(n).clone() for n in graph.outputs[:-1]]
            return (*feats, depth)
        return depth

    @staticmethod
    def backward(ctx, *grads):
        eng = ctx.eng
        if eng._fwd_id != ctx.fwd_id:
            raise RuntimeError("gdn_b200: backward() called after another forward of the same module and shape; the "
                               "engine keeps one set of activations (call backward before the next forward)")
        (depth,) = ctx.saved_tensors
        if any(g is not None for g in grads[:-1]):
            raise NotImplementedError("gdn_b200: gradients through the feature maps returned by forward(x, istrain=True) are "
                                      "not implemented in the module API (the reference never back-propagates through them; "
                                      "RtoDTrainStep(guidance_grad=True) is the supported guidance-gradient path)")
        dout = grads[-1]
        if dout is None:
            dout = torch.zeros_like(depth)
        dpre = dout * (1.0 - depth * depth)
        eng.flat_grad.zero_()
        eng.backward(dpre)
        out = [None, None, None, None, None]
        for n in ctx.names:
            out.append(eng.grad[n].clone() if n in eng.grad else None)
        return tuple(out)


def run_network(module, x, istrain):
    _refuse_replica(module)
    _require_cuda(x)
    graph = module.__dict__.get("_gdn_graph")
    if graph is None:
        graph = module.gdn_graph()
        module.__dict__["_gdn_graph"] = graph
    _engines(module)
    _check_norm(module)
    if x.dim() != 4 or x.shape[1] != graph.cin:
        raise ValueError("gdn_b200: expected input (N, %d, H, W), got %s" % (graph.cin, tuple(x.shape)))
    if x.shape[2] % 16 or x.shape[3] % 16:
        raise ValueError("gdn_b200: height and width must be multiples of 16 (got %dx%d), like the reference networks "
                         "themselves (SURVEY.md 0.4: odd sizes break the skip concatenation)" % (x.shape[2], x.shape[3]))
    named = [(n, p) for n, p in module.named_parameters()]
    wants_grad = torch.is_grad_enabled() and any(p.requires_grad for _, p in named)
    if torch.is_grad_enabled() and x.requires_grad:
        raise NotImplementedError("gdn_b200: input gradients through the module API are not implemented (nothing on the "
                                  "reference path needs d/d input); RtoDTrainStep(guidance_grad=True) back-propagates "
                                  "through the frozen DtoD encoder")
    if wants_grad and not module.training and not module.__dict__.get("_gdn_warned_eval_grad"):
        import warnings
        module.__dict__["_gdn_warned_eval_grad"] = True
        warnings.warn("gdn_b200: eval-mode forward with autograd enabled returns a tensor WITHOUT grad_fn (the eval engine "
                      "folds BatchNorm and keeps no activations); wrap evaluation in torch.no_grad() like the reference's "
                      "validate(), or call model.train() to get gradients", RuntimeWarning, stacklevel=3)
    needs_grad = module.training and wants_grad
    if needs_grad:
        names = tuple(n for n, _ in named)
        return _NetFn.apply(x, module, graph, bool(istrain is True), names, *[p for _, p in named])
    want = graph.outputs if istrain is True else ()
    eng = get_engine(module, graph, x, train=module.training, backward=False, want=want)
    with torch.no_grad():
        eng.forward(x)
        if module.training:
            _after_train_forward(module)
        depth = eng.depth().clone()
        if istrain is True:
            return tuple(eng.value_nchw(n).clone() for n in graph.outputs[:-1]) + (depth,)
        return depth


def encoder_features(module, x):
    """(x1, x2, x4, x6) of a frozen (eval-mode) network as fp32 NHWC buffers WITHOUT running the decoder: all the
    guidance loss consumes (trainer.py:700,703).  Buffers are valid until the next call with the same shape."""
    _require_cuda(x)
    graph = module.__dict__.get("_gdn_graph")
    if graph is None:
        graph = module.gdn_graph()
        module.__dict__["_gdn_graph"] = graph
    _engines(module)
    _check_norm(module)
    names = graph.encoder_outputs
    eng = get_engine(module, graph, x, train=module.training, backward=False, want=names, stop_after=names[-1])
    with torch.no_grad():
        eng.forward(x)
        if module.training:
            _after_train_forward(module)
    return [eng.value(n) for n in names]


# ----------------------------------------------------------------------------------------- stand-alone blocks
def _block_graph(block):
    cin, cout, k, pad, stride = block._cfg
    g = Graph(type(block).__name__, cin)
    if block._kind == "res":
        g.units.append(Unit("main.0", ("in",), "h", cin, cout, k, 1, pad, bn="main.1", relu=True))
        g.units.append(Unit("main.3", ("h",), "out", cout, cout, k, 1, pad, bn="main.4", resid="in"))
    elif block._kind == "conv":
        g.units.append(Unit("main.1", ("in",), "out", cin, cout, k, stride, pad, reflect=pad > 0, bn="main.2", relu=True))
    else:
        g.units.append(Unit("main.0", ("in",), "out", cin, cout, k, stride, pad, transposed=True, bn="main.1", relu=True))
    g.outputs = ("out",)
    return g


class _BlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, block, graph, names, *params):
        eng = get_engine(block, graph, x, train=True, backward=True, want=("out",))
        eng.forward(x)
        eng._fwd_id = getattr(eng, "_fwd_id", 0) + 1
        _after_train_forward(block)
        ctx.eng, ctx.fwd_id, ctx.names = eng, eng._fwd_id, names
        return eng.value_nchw("out").clone()

    @staticmethod
    def backward(ctx, dout):
        eng = ctx.eng
        if eng._fwd_id != ctx.fwd_id:
            raise RuntimeError("gdn_b200: backward() after another forward of the same block and shape")
        eng.flat_grad.zero_()
        eng.backward(dout_nhwc=dout.permute(0, 2, 3, 1))
        dx = eng.dact["in"].permute(0, 3, 1, 2).clone() if "in" in eng.dact else None
        out = [dx, None, None, None]
        for n in ctx.names:
            out.append(eng.grad[n].clone() if n in eng.grad else None)
        return tuple(out)


def run_block(block, x):
    _refuse_replica(block)
    _require_cuda(x)
    graph = block.__dict__.get("_gdn_graph")
    if graph is None:
        graph = _block_graph(block)
        block.__dict__["_gdn_graph"] = graph
    _engines(block)
    _check_norm(block)
    if graph.cin < 64 and x.requires_grad:
        raise NotImplementedError("gdn_b200: input gradients of thin-channel (first-layer) blocks are not needed by the "
                                  "reference path and not implemented")
    named = list(block.named_parameters())
    if block.training and torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for _, p in named)):
        names = tuple(n for n, _ in named)
        return _BlockFn.apply(x, block, graph, names, *[p for _, p in named])
    eng = get_engine(block, graph, x, train=block.training, backward=False, want=("out",))
    with torch.no_grad():
        eng.forward(x)
        if block.training:
            _after_train_forward(block)
        return eng.value_nchw("out").clone()
