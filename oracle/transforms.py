"""Restatement of the reference's input transforms (TEST INFRASTRUCTURE).

Reference: /root/reference/src/transform_list.py -- ArrayToTensor (:95-113) + Normalize (:84-93) verbatim arithmetic
(float / 255, then sub_(0.5).div_(0.5)); RandomHorizontalFlip (:161-169) = np.fliplr; RandomScaleCrop (:189-203) =
``imresize(im, (scaled_h, scaled_w))`` + crop.  The loader hands the transforms FLOAT32 arrays
(src/datasets/datasets_list.py:81-86 load_as_float = imread(..).astype(np.float32)), so scipy.misc.imresize first
byte-scales every image (per-image min-max stretch of the whole array to 0..255) and then resizes the 8-bit image with
PIL: that is oracle/imresize.py (resize pinned bit-for-bit on Pillow; bytescale restated from the published SciPy 1.2
source, unpinned -- scipy.misc no longer exists).
"""
import numpy as np
import torch


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
        from .imresize import imresize
        sh, sw, oy, ox = (int(v) for v in crop)
        h, w, c = im.shape
        arr = im.astype(np.float32)                                                     # load_as_float
        z = imresize(arr[:, :, 0] if c == 1 else arr, (sh, sw))                         # :197 (2-D 'L' / HxWx3 'RGB')
        if z.ndim == 2:
            z = z[:, :, None]
        im = z[oy:oy + h, ox:ox + w]                                                    # :201
    return im
