"""shared helpers for the test-suite"""
import contextlib
import io
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_SHAPES = {}


def build_module(name, **kw):
    """construct a product module quietly (the reference classes print '- norm : Batch')"""
    from gdn_pytorch_b200 import AE_model_unet as M
    with contextlib.redirect_stdout(io.StringIO()):
        return getattr(M, name)(**kw)


def shapes_of(name):
    if name not in _SHAPES:
        m = build_module(name, init_weights=False)
        _SHAPES[name] = {k: v.shape for k, v in m.state_dict().items()}
    return _SHAPES[name]


def golden(fname):
    return np.load(os.path.join(GOLDEN, fname))


def relerr(a, b):
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-20)


def engine_act_grad(eng, name, ref_grad, masks):
    """(engine value, reference value) of dL/d(tensor `name`) as NCHW fp32.  When the engine fused the BatchNorm-backward
    statistics of the tensor's producer into the last input-gradient convolution, the total gradient only exists as the
    bf16 buffer eng.gm[name] with the producer's ReLU mask already applied -- the reference is masked the same way."""
    if name in getattr(eng, "gm", {}):
        got = eng.gm[name].float().permute(0, 3, 1, 2)
        want = ref_grad * masks[name] if name in masks else ref_grad
        return got, want
    return eng.dact[name].permute(0, 3, 1, 2), ref_grad
