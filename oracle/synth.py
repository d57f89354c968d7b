"""Deterministic synthetic weights and inputs shared by the golden generator, tests and bench (TEST INFRA).

Weights are generated per key from (seed, crc32(key)) so a state_dict can be rebuilt anywhere from its
key->shape table alone; nothing large has to be committed.  Inputs follow SURVEY.md 8d:
rgb U[-1,1]; dense depth = vertical ramp + smooth blobs in [-1,1]; sparse = dense with 70 % of the pixels
set to exactly -1.0 (the `sparse > -1` validity test of trainer.py:706).
"""
import zlib
import math
import torch


def _gen(seed, key):
    g = torch.Generator()
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))
    return g


def synth_state_dict(shapes, seed=0, bn_random=True):
    """shapes: {key: torch.Size}.  Conv weights U(+-1/sqrt(fan_in)) like the reference's _initialize_weights
    (AE_model_unet.py:249-261); BN affine/running stats randomised so eval-mode folding is exercised."""
    sd = {}
    for k, shp in shapes.items():
        g = _gen(seed, k)
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.int64)
        elif len(shp) == 4:
            fan = shp[1] * shp[2] * shp[3]
            s = 1.0 / math.sqrt(fan)
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * s
        elif k.endswith("running_var"):
            sd[k] = (torch.rand(shp, generator=g) + 0.5) if bn_random else torch.ones(shp)
        elif k.endswith("running_mean"):
            sd[k] = (torch.rand(shp, generator=g) - 0.5) * 0.2 if bn_random else torch.zeros(shp)
        elif k.endswith("weight"):
            sd[k] = (torch.rand(shp, generator=g) + 0.5) if bn_random else torch.ones(shp)
        elif k.endswith("bias"):
            sd[k] = (torch.rand(shp, generator=g) - 0.5) * 0.4 if bn_random else torch.zeros(shp)
        else:
            raise KeyError(k)
    return sd


def synth_rgb(B, H, W, seed=0):
    g = torch.Generator().manual_seed(seed + 11)
    return torch.rand((B, 3, H, W), generator=g) * 2 - 1


def synth_depth(B, H, W, seed=0):
    g = torch.Generator().manual_seed(seed + 23)
    ys = torch.linspace(1.0, -1.0, H).view(1, 1, H, 1).expand(B, 1, H, W)
    low = torch.rand((B, 1, max(H // 8, 1), max(W // 8, 1)), generator=g) * 2 - 1
    blobs = torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=False)
    return (0.6 * ys + 0.5 * blobs).clamp(-1, 1).contiguous()


def synth_sparse(dense, seed=0, keep=0.3):
    g = torch.Generator().manual_seed(seed + 37)
    m = torch.rand(dense.shape, generator=g) < keep
    return torch.where(m, dense, torch.full_like(dense, -1.0))


def synth_pred(B, H, W, seed=0):
    g = torch.Generator().manual_seed(seed + 41)
    return torch.tanh(torch.randn((B, 1, H, W), generator=g))
