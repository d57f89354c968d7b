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
