"""GPU (-m gpu): TRAIN-mode (batch-statistics BatchNorm) numerics of the B200 path at the headline shape -- 128 x 416,
the networks of the bench step -- against the fp32 reference arithmetic (oracle restatement of the reference classes run
with torch / cuDNN on the same GPU, TF32 off) and against its bf16-operand emulation (oracle/model.py bf16=True).

Why a "warm" weight state: at random init a train-mode network amplifies ANY perturbation by one to two orders of
magnitude -- the bf16-operand emulation of the reference itself moves the fp32 output by 25 % (AutoEncoder_2) / 195 %
(AutoEncoder_DtoD), measured in profiles/r02b_parity_probe.log -- so no implementation with bf16 operands (north star
item 1) can be pinned there.  After 30 reference (fp32 oracle) training steps the same networks are well conditioned
(emulation vs fp32: 1.1e-2 / 2.1e-2) and the comparison is meaningful; the tests below run in that state and ALSO bound
the product by the emulation's own distance, so that a regression of the conditioning cannot hide a regression of the
kernels.  Measured values (tests/parity_probe.py, profiles/r02*_parity_*.log) are quoted next to every bound.
Reference: /root/reference/src/AE_model_unet.py:312-368, 527-574; src/trainer.py:696-768."""
import pytest
import torch

from tests import parity_probe as PP
from tests.util import relerr

pytestmark = pytest.mark.gpu
B, H, W = 4, 128, 416
_STATE = {}


def _conditioning(name, sd):
    """bf16-operand emulation of the reference vs the fp32 reference on the test input: how much ANY implementation with
    bf16 operands is expected to differ in this weight state"""
    from oracle import model as OM
    x = PP.inputs_for(name, B, H, W, 3)
    with torch.no_grad():
        r32 = OM.FORWARDS[name]({k: v.clone() for k, v in sd.items()}, x, istrain=False, train=True)
        r16 = OM.FORWARDS[name]({k: v.clone() for k, v in sd.items()}, x, istrain=False, train=True, bf16=True)
    return relerr(r16, r32)


def _warm(name):
    """the warm weight state, built once per network.  The precondition of every test below is that the state is well
    conditioned (emulation vs fp32 <= 3e-2); the warm-up is deterministic (parity_probe.no_tf32), and should a software
    stack land on a worse state it is trained further (10 more reference steps, at most 3 times) before the tests give up"""
    if name not in _STATE:
        PP.no_tf32()
        sd = PP.state(name, "warm", B, H, W)
        for _ in range(3):
            if _conditioning(name, sd) <= 3e-2:
                break
            sd = PP.warm_up(name, {k: v.detach().clone() for k, v in sd.items()}, B, H, W, steps=10)
        _STATE[name] = sd
    return {k: v.clone() for k, v in _STATE[name].items()}


#                       engine vs fp32 (measured)      emulation vs fp32 (measured)
FWD_BOUND = {"AutoEncoder_2": 2.0e-2,      # 1.00e-2                        1.07e-2
             "AutoEncoder_DtoD": 5.0e-2}   # 2.96e-2                        2.15e-2


@pytest.mark.parametrize("name", ["AutoEncoder_2", "AutoEncoder_DtoD"])
def test_train_mode_forward_at_headline_shape(name):
    """whole-network train-mode forward, B = 4, 128 x 416: depth max-rel error vs the fp32 reference arithmetic"""
    from oracle import model as OM
    PP.no_tf32()
    sd = _warm(name)
    x = PP.inputs_for(name, B, H, W, 3)
    with torch.no_grad():
        r32 = OM.FORWARDS[name]({k: v.clone() for k, v in sd.items()}, x, istrain=False, train=True)
        r16 = OM.FORWARDS[name]({k: v.clone() for k, v in sd.items()}, x, istrain=False, train=True, bf16=True)
        m = PP.product_module(name, sd, H, W)
        m.train()
        got = m(x, istrain=False)
    e32, e16, cond = relerr(got, r32), relerr(got, r16), relerr(r16, r32)
    print("%s train fwd: engine-vs-fp32 %.3e engine-vs-bf16emu %.3e bf16emu-vs-fp32 %.3e" % (name, e32, e16, cond))
    assert cond <= 4e-2, "weight state is not well conditioned (emulation vs fp32 %.3e)" % cond
    assert e32 <= FWD_BOUND[name], (e32, cond)
    assert e32 <= 1.6 * cond + 5e-3, (e32, cond)       # no further from fp32 than bf16 operands alone explain
    assert e16 <= 1.6 * cond + 5e-3, (e16, cond)


@pytest.mark.parametrize("name", ["AutoEncoder_2", "AutoEncoder_DtoD"])
def test_parameter_gradients_at_headline_shape(name):
    """module API (model.train(); loss.backward(), trainer.py:466-468) vs fp32 autograd through the reference arithmetic:
    whole gradient cosine >= 0.998 / L2 <= 8e-2 (measured 0.99935 / 4.8e-2 and 0.99985 / 1.8e-2); per parameter tensor
    cosine >= 0.99 for every tensor that carries a measurable share of the gradient (BatchNorm betas in front of another
    batch-statistics BatchNorm have a gradient that is zero up to border effects -- pure noise in ANY implementation:
    the emulation's cosine to fp32 is negative there too -- so tensors are weighted by their share of |g|^2)."""
    from oracle import synth
    PP.no_tf32()
    sd = _warm(name)
    x = PP.inputs_for(name, B, H, W, 3)
    tgt = synth.synth_depth(B, H, W, 7).to(PP.dev)
    g32, _ = PP.oracle_grads(name, sd, x, tgt, False)
    m = PP.product_module(name, sd, H, W)
    m.train()
    out = m(x, istrain=False)
    loss = ((out - tgt) ** 2).mean()
    m.zero_grad()
    loss.backward()
    got = {k: p.grad.detach().double().flatten() for k, p in m.named_parameters()}
    ref = {k: g32[k].double().flatten() for k in got}
    a, b = torch.cat(list(got.values())), torch.cat([ref[k] for k in got])
    cos = (torch.dot(a, b) / (a.norm() * b.norm())).item()
    l2 = ((a - b).norm() / b.norm()).item()
    total2 = (b.norm() ** 2).item()
    bad_energy, worst, worst_small = 0.0, (1.0, None), (1.0, None)
    for k in got:
        nb = ref[k].norm().item()
        if nb == 0.0:
            continue
        c = (torch.dot(got[k], ref[k]) / (got[k].norm() * nb + 1e-300)).item()
        if c < 0.99:
            bad_energy += nb * nb
        if nb * nb >= 1e-2 * total2 and c < worst[0]:
            worst = (c, k)
        if nb * nb >= 1e-3 * total2 and c < worst_small[0]:
            worst_small = (c, k)
    print("%s grads: whole cos %.5f L2 %.3e; tensors below cos 0.99 carry %.2e of |g|^2; worst tensor with >= 1 %% of |g|^2 %s, "
          "with >= 0.1 %% %s" % (name, cos, l2, bad_energy / total2, worst, worst_small))
    assert cos >= 0.998 and l2 <= 8e-2, (cos, l2)
    assert worst[0] >= 0.99, worst                      # every tensor with >= 1 % of the gradient energy
    # 0.1 % .. 1 %: per-channel sums over ~10^6 pixels of signed gradients (BatchNorm betas) cancel heavily, so the bf16
    # rounding of the gradient operands shows: measured 0.975 for res64_up1.main.4.bias of AutoEncoder_DtoD (64 values)
    assert worst_small[0] >= 0.95, worst_small
    assert bad_energy <= 1e-2 * total2, bad_energy / total2      # measured 1.6e-3 .. 3.6e-3: many small tensors


def test_rtod_loss_trajectory_50_steps():
    """the fused RtoD step (trainer.py:696-768) for 50 steps from the reference's own init (gamma 1, beta 0), lr 2e-5
    (option.py:18), 4 alternating batches, against the same step with torch ops in fp32 AND against the bf16-operand
    emulation of the reference on the same run.  Training trajectories separate with time, and how fast depends on the
    fp32 reference's own rounding (cuDNN algorithm choice): measured worst deviations of (product, emulation) from fp32
    over the 50 steps were (1.6 %, 1.0 %) with cuDNN autotuned algorithms (profiles/r02c_parity_traj.log) and
    (3.1 %, 3.4 %) with the deterministic ones (profiles/r02l_pytest.log).  So the bound is relative to what bf16 operands
    alone do to the reference on this very run: every loss term of the product stays within 2.5 % of fp32, or within
    1.25 x the emulation's own deviation + 0.5 %, whichever is larger -- and never beyond 6 %."""
    worst, emu = PP.probe_traj(B, H, W, steps=50, lr=2e-5, kind="init")
    for k, v in worst.items():
        assert v <= max(2.5e-2, 1.25 * emu[k] + 5e-3), (k, v, emu[k])
        assert v <= 6e-2, (k, v)
