"""Writes tests/golden/imresize.npz from Pillow's own resize (TEST INFRASTRUCTURE; run in the dev container:
``python -m oracle.gen_golden_imresize``).  Pillow is the third-party library scipy.misc.imresize delegates to
(/root/reference/src/depth_extract.py:23-58); the fixtures pin oracle/imresize.py::resize_u8 -- and through it the
device kernels -- where Pillow is not importable.  Inputs are regenerated from seeds; of PIL's outputs a SHA-256 and a 24x24 corner are stored (small fixtures)."""
import os

import numpy as np

CASES = [  # (seed, h, w, c, oh, ow)
    (0, 375, 1242, 3, 128, 416),   # KITTI frame -> network size (down, both axes)
    (1, 128, 416, 1, 375, 1242),   # depth map -> original size (up, both axes)
    (2, 37, 53, 3, 128, 416),
    (3, 200, 300, 3, 100, 300),    # vertical pass only
    (4, 100, 50, 1, 33, 77),       # down in y, up in x
    (5, 480, 640, 3, 128, 416),    # NYU frame
]


def case_input(seed, h, w, c):
    img = np.random.RandomState(seed).randint(0, 256, (h, w, c)).astype(np.uint8)
    return img[:, :, 0] if c == 1 else img


def digest(a):
    import hashlib
    a = np.ascontiguousarray(a)
    return hashlib.sha256(repr(a.shape).encode() + a.tobytes()).hexdigest()


def main():
    from PIL import Image
    out = {}
    for i, (seed, h, w, c, oh, ow) in enumerate(CASES):
        img = case_input(seed, h, w, c)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
        out["case%d_sha256" % i] = np.array(digest(ref))           # whole output, bit-exact
        out["case%d_corner" % i] = ref[:24, :24].copy()            # a readable piece of it
    import PIL
    out["pillow_version"] = np.array(PIL.__version__)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "imresize.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
