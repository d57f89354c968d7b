"""Restatement of the reference optimizer step (TEST INFRASTRUCTURE).

Reference: optim.Adam(params, lr, [0.9, 0.999], eps=1e-8, weight_decay=5e-4) at
/root/reference/src/GDN_main.py:157,173,184 -- coupled L2 decay (not AdamW), bias-corrected,
denominator sqrt(v_hat) + eps in torch's formulation: sqrt(v)/sqrt(1-b2^t) + eps.
"""
import math
import torch


def adam_step(p, g, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8, wd=5e-4):
    """in-place on p, m, v (all fp32 tensors); ``step`` is the 1-based step count"""
    g = g + wd * p
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)
    return p
