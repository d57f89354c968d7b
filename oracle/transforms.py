"""Restatement of the reference's input transforms (TEST INFRASTRUCTURE).

Reference: /root/reference/src/transform_list.py -- ArrayToTensor (:95-113) + Normalize (:84-93) verbatim arithmetic
(float / 255, then sub_(0.5).div_(0.5)); RandomHorizontalFlip (:161-169) = np.fliplr; RandomScaleCrop (:189-203) =
imresize to (scaled_h, scaled_w) + crop.  scipy.misc.imresize (PIL bilinear, uint8 result) no longer exists in SciPy
(SURVEY.md Appendix C), so the zoom is restated as pixel-centre-aligned bilinear interpolation rounded to uint8 --
parity for that step is therefore "unpinned" against the reference and held to +-1 grey level.
"""
import numpy as np
import torch
import torch.nn.functional as F


def to_tensor_normalize(img_u8):
    """(H, W, C) or (H, W) uint8 numpy -> (C, H, W) fp32 in [-1, 1]"""
    im = np.asarray(img_u8)
    if im.ndim == 2:
        im = im[:, :, None]
    t = torch.from_numpy(np.ascontiguousarray(im.transpose(2, 0, 1))).float() / 255     # :112
    for ch in t:
        ch.sub_(0.5).div_(0.5)                                                          # :92
    return t


def flip_scale_crop(img_u8, flip, crop):
    """crop = (scaled_h, scaled_w, off_y, off_x) or None; returns uint8 (H, W, C)"""
    im = np.asarray(img_u8)
    if im.ndim == 2:
        im = im[:, :, None]
    if flip:
        im = np.copy(np.fliplr(im))                                                     # :166
    if crop is not None:
        sh, sw, oy, ox = (int(v) for v in crop)
        h, w = im.shape[0], im.shape[1]
        t = torch.from_numpy(np.ascontiguousarray(im.transpose(2, 0, 1))).float()[None]
        z = F.interpolate(t, size=(sh, sw), mode="bilinear", align_corners=False)[0]
        z = z.round().clamp(0, 255).byte().numpy().transpose(1, 2, 0)
        im = z[oy:oy + h, ox:ox + w]                                                    # :201
    return im
