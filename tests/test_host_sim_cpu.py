"""CPU: kernels whose index arithmetic is written as __host__ __device__ pieces are run thread by thread on the host
(tests/host/*.cu, test-only shared objects built here with nvcc) and compared with the oracle -- a way to check a new
elementwise kernel's addressing without a GPU.  The -m gpu tests check the real kernels; the product has no host path."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "host", "_build")


def _build(name):
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(ROOT, "tests", "host", name + ".cu")
    out = os.path.join(BUILD, "lib" + name + ".so")
    hdrs = [os.path.join(ROOT, "gdn_pytorch_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "gdn_pytorch_b200", "csrc"))
            if f.endswith(".cuh")]
    if not os.path.isfile(out) or os.path.getmtime(out) < max(os.path.getmtime(p) for p in [src] + hdrs):
        subprocess.run(["nvcc", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-o", out, src], check=True)
    return C.CDLL(out)


FOLD_CASES = [  # (N, H, W, C, pad, reflect, up, dilate, accumulate, ctot, c_off)
    (2, 8, 12, 64, 1, 1, 1, 0, 0, 64, 0),      # upconv: x2 bilinear + reflection border 1
    (1, 6, 10, 128, 3, 1, 1, 0, 1, 128, 0),    # border 3 (k7), accumulate
    (2, 5, 7, 64, 2, 1, 1, 0, 0, 64, 0),
    (2, 8, 12, 64, 1, 1, 0, 0, 0, 64, 0),      # stride-2 reflect conv: reflection only
    (1, 9, 11, 16, 3, 1, 0, 0, 1, 16, 0),
    (2, 6, 8, 64, 0, 0, 0, 1, 0, 64, 0),       # transposed conv: zero dilation
    (1, 4, 6, 64, 2, 0, 1, 0, 0, 64, 0),       # upsample + zero border
    (2, 7, 9, 32, 1, 1, 0, 0, 0, 96, 32),      # channel slice of a wider buffer (virtual concat source)
]


@pytest.mark.parametrize("case", FOLD_CASES, ids=lambda c: "x".join(map(str, c)))
def test_fold_rows2_host_simulation_is_the_adjoint(case):
    N, H, W, Cc, P, reflect, up, dilate, acc, ctot, c_off = case
    lib = _build("fold_sim")
    g = torch.Generator().manual_seed(sum(case))
    sc = 2 if (up or dilate) else 1
    dpad = torch.randn((N, H * sc + 2 * P, W * sc + 2 * P, ctot), generator=g)
    base = torch.randn((N, H, W, Cc), generator=g)
    out = base.clone()
    rc = lib.fold_rows2_host(C.c_void_p(dpad.data_ptr()), ctot, c_off, N, H, W, Cc, P, reflect, up, dilate,
                             C.c_void_p(out.data_ptr()), acc, 96, 5)       # odd thread / CTA counts on purpose
    assert rc == 0
    x = torch.zeros((N, Cc, H, W), dtype=torch.float64, requires_grad=True)
    z = x
    if up:
        z = F.interpolate(z, scale_factor=2, mode="bilinear", align_corners=False)
    elif dilate:
        zz = torch.zeros((N, Cc, 2 * H, 2 * W), dtype=torch.float64)
        zz[:, :, ::2, ::2] = z
        z = zz
    if P:
        z = F.pad(z, (P,) * 4, mode="reflect" if reflect else "constant")
    (gx,) = torch.autograd.grad(z, x, dpad[..., c_off:c_off + Cc].double().permute(0, 3, 1, 2))
    ref = gx.permute(0, 2, 3, 1) + (base.double() if acc else 0)
    assert torch.allclose(out.double(), ref, rtol=1e-5, atol=1e-5), (out.double() - ref).abs().max()


PACK_CASES = [  # (cout, cin, k, transposed, a_pad_extra, scaled)
    (64, 64, 9, False, 0, False), (128, 64, 7, False, 0, True), (256, 128, 5, False, 0, False),
    (512, 256, 3, False, 0, True), (512, 512, 3, False, 0, False), (512, 1024, 1, False, 0, True),
    (128, 64, 4, False, 0, False), (64, 128, 4, True, 0, False), (1, 64, 9, False, 15, False), (40, 24, 3, False, 8, True),
]


@pytest.mark.parametrize("case", PACK_CASES, ids=lambda c: "x".join(map(str, c)))
@pytest.mark.parametrize("kind", ["forward", "dgrad"])
def test_pack_v2_host_simulation_matches_the_layout(case, kind):
    """forward pack [tap][cout][cin] and input-gradient pack [tap][cin][cout] (flipped taps for Conv2d) of Conv2d /
    ConvTranspose2d weights, with zero padding of a and per-a scaling -- the descriptors engine.py builds"""
    cout, cin, kk_, transposed, apad_extra, scaled = case
    lib = _build("pack_sim")
    lib.pack_v2_host.restype = C.c_int
    g = torch.Generator().manual_seed(cout + cin + kk_)
    T = kk_ * kk_
    w = torch.randn((cin, cout, kk_, kk_) if transposed else (cout, cin, kk_, kk_), generator=g)
    if kind == "forward":
        A, B = cout, cin
        sa, sb = (T, cout * T) if transposed else (cin * T, T)
        flip = 1 if transposed else 0
        ref = (w.permute(2, 3, 1, 0) if transposed else w.permute(2, 3, 0, 1)).reshape(T, cout, cin)
    else:
        A, B = cin, cout
        sa, sb = (cout * T, T) if transposed else (T, cin * T)
        flip = 0 if transposed else 1
        ref = (w.permute(2, 3, 0, 1) if transposed else w.permute(2, 3, 1, 0)).reshape(T, cin, cout)
    if flip:
        ref = ref.flip(0)
    Apad, Bpad = A + apad_extra, B
    scale = torch.randn(A, generator=g) if scaled else None
    if scale is not None:
        ref = ref * scale[None, :, None]
    full = torch.zeros((T, Apad, Bpad))
    full[:, :A, :B] = ref
    out = torch.full((T, Apad, Bpad), 7.0).to(torch.bfloat16)
    tb = lib.pack_v2_host(kk_, kk_, A, B, Apad, Bpad, C.c_longlong(sa), C.c_longlong(sb), C.c_longlong(kk_), C.c_longlong(1), flip,
                          C.c_void_p(w.data_ptr()), C.c_void_p(scale.data_ptr() if scale is not None else None),
                          C.c_void_p(out.data_ptr()), 96)
    assert tb in (16, 32, 64), tb
    assert torch.equal(out, full.to(torch.bfloat16))


def test_lane_partitioned_radix_select_equals_the_sequential_walk():
    """metrics_select_state (csrc/loss_metrics.cu) finds the histogram bin that holds rank k with one warp: lane l owns
    bins 8l .. 8l+7, an inclusive scan over the lanes finds the first lane whose cumulative count exceeds k, that lane walks
    its 8 bins (the last one untested).  This restates that partition in numpy and checks it against the sequential walk
    over bins 0 .. 254 it replaced (bin 255 is the fall-through of both), incl. empty histograms, single-bin histograms,
    everything in bin 255 and ranks beyond the total."""
    def sequential(hist, k):
        b = 0
        while b < 255:
            c = int(hist[b])
            if k < c:
                break
            k -= c
            b += 1
        return b, k

    def by_lanes(hist, k):
        c = hist.reshape(32, 8).astype(np.int64)
        tot = c.sum(1)
        incl = np.cumsum(tot)
        hit = k < incl
        if not hit.any():
            return 255, k - (int(incl[31]) - int(c[31, 7]))
        lane = int(np.argmax(hit))
        myk = k - int(incl[lane] - tot[lane])
        j = 0
        while j < 7:
            if myk < c[lane, j]:
                break
            myk -= int(c[lane, j])
            j += 1
        return lane * 8 + j, myk

    rng = np.random.RandomState(0)
    for t in range(20000):
        mode = t % 5
        h = rng.randint(0, 50, 256)
        if mode == 1:
            h[rng.randint(0, 256, 200)] = 0
        elif mode == 2:
            h[:] = 0
            h[rng.randint(0, 256)] = rng.randint(1, 1000)
        elif mode == 3:
            h[:255] = 0
            h[255] = rng.randint(0, 5)
        n = int(h.sum())
        k = int(rng.randint(0, max(n, 1) + (3 if mode == 4 else 0))) if n > 0 else 0
        assert sequential(h, k) == by_lanes(h, k), (mode, k)
