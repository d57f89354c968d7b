"""numpy restatement of the demo path's image resizing (TEST INFRASTRUCTURE -- never imported by the product).

/root/reference/src/depth_extract.py:23-58 (class Resize) calls ``scipy.misc.imresize(img, size, 'bilinear')`` on the
float32 input image (:86) and on the float network output (:138).  scipy.misc.imresize is a THIRD-PARTY function
that is absent here (removed in SciPy 1.3; this container has SciPy 1.18; the reference pins no version, its README
era is SciPy 1.0-1.2).  Its published algorithm (scipy/misc/pilutil.py of SciPy 1.2.x), restated:

    imresize(arr, size, 'bilinear') = fromimage(toimage(arr).resize((w, h), resample=PIL.Image.BILINEAR))
    toimage(arr)  = PIL image ('L' for 2-D, 'RGB' for HxWx3) of bytescale(arr)
    bytescale(a)  = a if a.dtype == uint8 else
                    uint8(clip((a - a.min()) * (255 / (a.max() - a.min() or 1)), 0, 255) + 0.5)

and PIL's BILINEAR resize of 8-bit images (Pillow src/libImaging/Resample.c): separable, horizontal pass then
vertical pass through an 8-bit intermediate image, triangle filter whose support is scaled by the down-scaling
factor (antialiasing), coefficients normalised in double precision and quantised to 22-bit fixed point, each output
= (2^21 + sum coef * pixel) >> 22 clamped to [0, 255].

Pinning: Pillow IS importable in the development container (12.2), so ``resize_u8`` is checked bit-for-bit against
``PIL.Image.resize`` (tests/test_oracle_golden.py::test_imresize_*, live when PIL imports, and against
tests/golden/imresize.npz written by oracle/gen_golden_imresize.py from PIL).  ``bytescale`` has no implementation
to run here: its arithmetic follows the published source (float32 arrays: float32 intermediates with the scale
rounded to float32 -- NumPy's legacy scalar promotion; float64 arrays, which is what the demo's output path builds
at :135: double) -- parity of that step is UNPINNED and says so.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bytescale(data):
    """scipy.misc.bytescale(data) with the defaults imresize/toimage use (cmin/cmax = data min/max, 0..255).
    float32 arrays: float32 arithmetic, the float64 scale scalar is rounded to float32 (NumPy's legacy promotion of
    array * scalar, the NumPy of the reference's era); float64 arrays: double arithmetic."""
    data = np.asarray(data)
    if data.dtype == np.uint8:
        return data
    ft = np.float64 if data.dtype == np.float64 else np.float32
    d = data.astype(ft)
    cmin, cmax = ft(d.min()), ft(d.max())
    cscale = ft(cmax - cmin)
    if cscale == 0:
        cscale = ft(1)
    scale = ft(255.0 / float(cscale))
    b = (d - cmin) * scale
    return (np.clip(b, ft(0), ft(255)) + ft(0.5)).astype(np.uint8)


def _coeffs(in_size, out_size):
    """precompute_coeffs + normalize_coeffs_8bpc of Resample.c for the triangle (bilinear) filter, box = whole image"""
    scale = float(np.float32(in_size) - np.float32(0)) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = np.zeros(ksize, dtype=np.float64)
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            w = 1.0 - a if a < 1.0 else 0.0
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        for x in range(ksize):
            kk[xx, x] = int(-0.5 + k[x] * (1 << PRECISION_BITS)) if k[x] < 0 else int(0.5 + k[x] * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, out_size, axis):
    """one separable pass along `axis` (0 = vertical, 1 = horizontal) of an (H, W, C) uint8 image"""
    in_size = img.shape[axis]
    bounds, kk = _coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.zeros((out_size,) + src.shape[1:], dtype=np.int64)
    for xx in range(out_size):
        xmin, n = bounds[xx]
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(n):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out.astype(np.uint8), 0, axis)


def resize_u8(img, size):
    """PIL.Image.fromarray(img).resize((size[1], size[0]), PIL.Image.BILINEAR) for uint8 (H, W) / (H, W, C) arrays;
    size = (out_h, out_w).  Horizontal pass first, then vertical; a pass whose size does not change is skipped."""
    img = np.asarray(img)
    assert img.dtype == np.uint8
    squeeze = img.ndim == 2
    if squeeze:
        img = img[:, :, None]
    oh, ow = size
    if ow != img.shape[1]:
        img = _pass(img, ow, 1)
    if oh != img.shape[0]:
        img = _pass(img, oh, 0)
    return img[:, :, 0] if squeeze else img


def imresize(arr, size):
    """scipy.misc.imresize(arr, size, 'bilinear') for 2-D and HxWx3 arrays, size = (h, w) -> uint8"""
    return resize_u8(bytescale(arr), size)


def demo_preprocess(img, size=(128, 416)):
    """depth_extract.py:84-91: resize -> CHW -> /255 -> (x - 0.5) / 0.5; img: (H, W, 3) float32/uint8 -> (1,3,h,w) fp32"""
    r = imresize(np.asarray(img, dtype=np.float32), size)
    t = r.transpose(2, 0, 1).astype(np.float32)[None]
    return ((t / np.float32(255)) - np.float32(0.5)) / np.float32(0.5)


def demo_postprocess(depth, org_size):
    """depth_extract.py:127-141: (128, 416) float depth -> uint8 (org_H, org_W) image as written by imsave
    (the F.interpolate to (128, 416) at :127 is the identity for a 128x416 output)"""
    img_ = np.empty(np.asarray(depth).shape)              # float64, like np.empty([128, 416]) at :135
    img_[...] = np.asarray(depth, dtype=np.float32)
    return imresize(img_, org_size)
