"""CPU: the engine's launch PLANS executed against a CPU emulation of the C ABI (tests/abi_emulator.py).

The engine is a plan builder (descriptors -> library calls); the plan's correctness is independent of the GPU.
These tests run it on raw pointers into CPU tensors, with every entry point emulated from the semantics documented
in include/gdn_b200.h, and compare with the oracle.  They do NOT exercise the CUDA kernels (tests -m gpu do) and
the product has no such path: gdn_pytorch_b200 refuses CPU tensors (tests/test_abi_cpu.py)."""
import pytest
import torch

from tests.abi_emulator import emulated_abi, engine_forward, run_ops
from tests.util import build_module, shapes_of, relerr, engine_act_grad

H, W, B = 32, 64, 2


def _module(name, seed, signed=False):
    from oracle import synth
    m = build_module(name, init_weights=False, height=H, width=W)
    sd = synth.synth_state_dict(shapes_of(name), seed=seed)
    if signed:
        for k in sd:
            if k.endswith(".weight") and sd[k].dim() == 1:
                sd[k][::3] *= -1.0
    m.load_state_dict(sd)
    m.eval()
    return m, sd


def _nchw(t):
    return t.permute(0, 3, 1, 2)


@pytest.mark.parametrize("name,cin", [("AutoEncoder_2", 3), ("AutoEncoder_DtoD", 1), ("AutoEncoder", 3)])
def test_eval_forward_plan_matches_oracle(name, cin):
    from gdn_pytorch_b200.engine import Engine
    from gdn_pytorch_b200.module_runtime import _params
    from oracle import model as OM, synth
    m, sd = _module(name, 0)
    x = synth.synth_rgb(B, H, W, 0) if cin == 3 else synth.synth_depth(B, H, W, 0)
    with emulated_abi() as emu:
        g = m.gdn_graph()
        eng = Engine(g, _params(m), B, H, W, train=False, want=g.outputs, device=torch.device("cpu"))
        engine_forward(eng, x)
        ref = OM.FORWARDS[name](sd, x, istrain=True)
        for nm, r in zip(g.outputs, ref):
            got = _nchw(eng.value(nm)).reshape(r.shape)
            assert relerr(got, r) <= 1e-2, (nm, relerr(got, r))
        assert emu.calls["gdn_conv2d"] == len(eng.units)
        assert emu.calls["gdn_head_gather"] == 1          # the 64 -> 1 head runs as taps-as-N 1x1 conv + shifted sum


def _l2rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def test_frozen_backward_plan_matches_autograd():
    """Engine(train=False, backward=True, input_grad=True, grad_seeds=...): the opt-in guidance gradient
     through the frozen DtoD encoder, against torch autograd through the oracle"""
    from gdn_pytorch_b200.engine import Engine
    from gdn_pytorch_b200.module_runtime import _params
    from gdn_pytorch_b200.ops import LossKernels
    from oracle import model as OM, losses as OL, synth
    m, sd = _module("AutoEncoder_DtoD", 1, signed=True)
    x, tar_in = synth.synth_depth(B, H, W, 3), synth.synth_depth(B, H, W, 4)
    with torch.no_grad():
        rt = OM.autoencoder_dtod(sd, tar_in, encoder_only=True)
    xr = x.clone().requires_grad_(True)
    lat = OL.latent_loss(OM.autoencoder_dtod(sd, xr, encoder_only=True), rt, with_grad=True)
    (ref,) = torch.autograd.grad(lat, xr)
    with emulated_abi():
        g = m.gdn_graph()
        names = g.encoder_outputs
        eng = Engine(g, _params(m), B, H, W, train=False, backward=True, want=names, stop_after=names[-1],
                     device=torch.device("cpu"), input_grad=True, grad_seeds=names)
        kern = LossKernels.__new__(LossKernels)
        kern.L = eng.L
        ft_tar = [r.permute(0, 2, 3, 1).contiguous() for r in rt]
        import gdn_pytorch_b200._lib as _l
        orig = _l.stream_ptr
        _l.stream_ptr = lambda: None
        try:
            for rep in range(2):            # seeds are rewritten every pass; nothing accumulates across passes
                engine_forward(eng, x)
                ft = [eng.value(n) for n in names]
                kern.latent_grad(ft, ft_tar, [eng.dact[n] for n in names])
                run_ops(eng.bwd)
        finally:
            _l.stream_ptr = orig
        got = eng.dact["in"].reshape(B, 1, H, W)
    assert torch.isfinite(got).all()
    cos = (torch.dot(got.flatten(), ref.flatten()) / (got.norm() * ref.norm())).item()
    assert cos >= 0.998, cos
    assert _l2rel(got, ref) <= 5e-2, _l2rel(got, ref)


@pytest.mark.parametrize("gname", ["mini_rtod", "mini_dtod", "mini_deep512"])
def test_training_plan_forward_backward_matches_autograd(gname):
    """The TRAINING plan (batch-statistics BatchNorm, BN backward, sub-pixel stride-2 dgrad, upsample / reflection /
    dilation folds, virtual concat, im2col'd thin layers and heads, weight-gradient unpack) on shallow graphs that
    contain every unit type -- same comparison as tests/test_gpu_network.py::test_backward_on_shallow_graphs, with the
    emulated ABI instead of the GPU: fp64 autograd with the engine's ReLU masks imposed."""
    from gdn_pytorch_b200.engine import Engine
    from oracle import synth
    from oracle.graph_interp import run_graph
    from tests import minigraphs
    g = getattr(minigraphs, gname)()
    sd = minigraphs.synth_params(g, 0)
    for k, v in sd.items():
        if not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
    x = synth.synth_rgb(B, H, W, 1) if g.cin == 3 else synth.synth_depth(B, H, W, 1)
    R = torch.rand((B, 1, H, W), generator=torch.Generator().manual_seed(5)) - 0.5
    names = [u.out for u in g.units]
    with emulated_abi() as emu:
        eng = Engine(g, sd, B, H, W, train=True, backward=True, want=names, device=torch.device("cpu"))
        with torch.no_grad():
            engine_forward(eng, x)
            # BatchNorm finalisation rides on the tail of the convolutions (gdn_conv_desc.fin_*): no separate launches
            assert "gdn_bn_finalize" not in emu.calls and all(cu.fin_fused for cu in eng.cu.values() if hasattr(cu, "fin"))
            masks = {u.out: (eng.value_nchw(u.out) > 0) for u in g.units if u.relu and not u.resid}
            out = eng.depth()
            eng.flat_grad.zero_()
            eng.dpre.copy_((R * (1 - out * out)).view(B, H, W))
            run_ops(eng.bwd)
    sd64 = {k: v.detach().double().requires_grad_(v.requires_grad) for k, v in sd.items()}
    T = run_graph(g, sd64, x.double(), train=True, relu_masks=masks)
    for t in T.values():
        if t.requires_grad:
            t.retain_grad()
    (T["out"] * R.double()).sum().backward()
    for n in names:
        assert relerr(eng.value_nchw(n), T[n].float()) <= 8e-2, n
    for u in g.units:
        if u.out != "out" and T[u.out].grad is not None:
            a, b = engine_act_grad(eng, u.out, T[u.out].grad.float(), masks)
            assert _l2rel(a, b) <= 3e-2, u.out
    checked = 0
    for k, v in sd64.items():
        if v.grad is None or v.grad.abs().max().item() < 1e-9:
            continue
        assert _l2rel(eng.grad[k], v.grad.float()) <= 3e-2, k
        checked += 1
    assert checked >= 2 * len(g.units) - 2
    # most BatchNorm-backward reductions ride on the epilogue of the convolution that completes the gradient
    assert len(eng._fused) >= 2 and len(eng.gm) >= 1, (sorted(eng._fused), sorted(eng.gm))


def test_instance_norm_branch_eval_is_running_stat_normalisation():
    """norm != 'Batch' (reference AE_model_unet.py:70-76, 88-93): nn.InstanceNorm2d(affine=True,
    track_running_stats=True) in eval mode == eval-mode BatchNorm on the same state_dict -- what lets the engine fold it"""
    g = torch.Generator().manual_seed(0)
    inn = torch.nn.InstanceNorm2d(8, affine=True, track_running_stats=True)
    bn = torch.nn.BatchNorm2d(8, affine=True, track_running_stats=True)
    with torch.no_grad():
        inn.weight.copy_(torch.rand(8, generator=g) + 0.5)
        inn.bias.copy_(torch.rand(8, generator=g) - 0.5)
        inn.running_mean.copy_(torch.rand(8, generator=g) - 0.5)
        inn.running_var.copy_(torch.rand(8, generator=g) + 0.5)
    bn.load_state_dict(inn.state_dict())
    x = torch.randn((3, 8, 5, 7), generator=g)
    assert torch.allclose(inn.eval()(x), bn.eval()(x), atol=1e-6)
    assert not torch.allclose(inn.train()(x), bn.eval()(x), atol=1e-3)      # train mode: per-sample statistics


def test_instance_norm_autoencoder_inference_plan(monkeypatch):
    """--norm Instance reaches nn.InstanceNorm2d only in the ``AutoEncoder`` class (reference AE_model_unet.py:146-154:
    AutoEncoder_2 / _DtoD print the choice but build their blocks with the default) -- the inference-only class of the
    live code.  Same state_dict keys; eval-mode inference through the engine (emulated ABI) against the UNMODIFIED
    reference class when /root/reference is present, else the oracle; train mode is refused (per-sample statistics are
    not implemented -- no silent substitution)."""
    import contextlib, io
    from gdn_pytorch_b200.engine import Engine
    from gdn_pytorch_b200.module_runtime import _params, _check_norm
    from oracle import model as OM, synth
    from oracle.refimport import reference_available, load_reference
    m = build_module("AutoEncoder", norm="Instance", init_weights=False, height=H, width=W)
    assert sum(isinstance(t, torch.nn.InstanceNorm2d) for t in m.modules()) == 7
    assert not any(isinstance(t, torch.nn.InstanceNorm2d)
                   for t in build_module("AutoEncoder_2", norm="Instance", init_weights=False).modules())
    sd = synth.synth_state_dict({k: v.shape for k, v in m.state_dict().items()}, seed=5)
    m.load_state_dict(sd)
    x = synth.synth_rgb(B, H, W, 2)
    m.train()
    with pytest.raises(NotImplementedError):
        _check_norm(m)
    m.eval()
    _check_norm(m)
    if reference_available():
        ae = load_reference()[0]
        with contextlib.redirect_stdout(io.StringIO()):
            ref = ae.AutoEncoder(norm="Instance", height=H, width=W)
        assert list(ref.state_dict().keys()) == list(m.state_dict().keys())
        ref.load_state_dict({k: v.clone() for k, v in sd.items()})
        ref.eval()
        monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)     # AutoEncoder.forward: x.cuda() (:161)
        with torch.no_grad():
            want = ref(x, istrain=False)
        monkeypatch.undo()
    else:
        want = OM.autoencoder(sd, x, istrain=False)
    with emulated_abi():
        g = m.gdn_graph()
        eng = Engine(g, _params(m), B, H, W, train=False, device=torch.device("cpu"))
        engine_forward(eng, x)
        got = eng.depth()
    assert relerr(got, want) <= 1e-2, relerr(got, want)


@pytest.mark.parametrize("switch", ["deterministic", "separate_finalize"])
def test_training_plan_switches(switch, monkeypatch):
    """two run-time switches of the training plan, on the emulated ABI: (a) GDN_DETERMINISTIC=1 -- the weight gradients go
    through the slab workspace (gdn_wgrad_desc.slabs, gdn_unpack_wgrad_slabs) and no fill precedes them; (b)
    GDN_FUSE_BNFIN=0 -- one gdn_bn_finalize launch per BatchNorm layer instead of the finalisation in the convolution tail.
    Both must produce the parameter gradients of the default plan."""
    from gdn_pytorch_b200.engine import Engine
    from oracle import synth
    from tests import minigraphs
    from tests.abi_emulator import EmulatedLib

    def grads(det, fuse_fin):
        g = minigraphs.mini_rtod()
        sd = minigraphs.synth_params(g, 0)
        for k, v in sd.items():
            if not k.endswith(("running_mean", "running_var")):
                v.requires_grad_(True)
        x = synth.synth_rgb(B, H, W, 1)
        R = torch.rand((B, 1, H, W), generator=torch.Generator().manual_seed(5)) - 0.5
        monkeypatch.setattr(EmulatedLib, "deterministic", det)
        monkeypatch.setenv("GDN_FUSE_BNFIN", "1" if fuse_fin else "0")
        with emulated_abi() as emu:
            eng = Engine(g, sd, B, H, W, train=True, backward=True, want=[u.out for u in g.units], device=torch.device("cpu"))
            with torch.no_grad():
                engine_forward(eng, x)
                out = eng.depth()
                eng.flat_grad.zero_()
                eng.dpre.copy_((R * (1 - out * out)).view(B, H, W))
                run_ops(eng.bwd)
            return {k: v.clone() for k, v in eng.grad.items()}, dict(emu.calls), eng

    ref, calls0, _ = grads(0, True)
    assert "gdn_bn_finalize" not in calls0 and "gdn_unpack_wgrad_slabs" not in calls0
    if switch == "deterministic":
        got, calls, eng = grads(1, True)
        assert eng.det and calls.get("gdn_unpack_wgrad_slabs", 0) == calls0["gdn_unpack_wgrad"] and "gdn_unpack_wgrad" not in calls
        tol = 1e-5            # the slab sums add the same products in another order
    else:
        got, calls, eng = grads(0, False)
        n_bn = sum(1 for u in eng.units if u.bn is not None)
        assert calls.get("gdn_bn_finalize", 0) == n_bn and not any(cu.fin_fused for cu in eng.cu.values() if hasattr(cu, "fin"))
        tol = 0.0             # same arithmetic, other launch
    for k in ref:
        assert (got[k] - ref[k]).abs().max().item() <= tol * max(1.0, ref[k].abs().max().item()), k

