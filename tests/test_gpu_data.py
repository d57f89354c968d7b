"""GPU (-m gpu): device-side input pipeline against the restated reference transforms (oracle/transforms.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda"


def _batch(n, h, w, c, seed):
    rs = np.random.RandomState(seed)
    return rs.randint(0, 256, size=(n, h, w, c) if c else (n, h, w)).astype(np.uint8)


@pytest.mark.parametrize("n,h,w,c", [(4, 128, 416, 3), (3, 32, 64, 1), (2, 16, 24, 0), (20, 128, 416, 3)])
def test_to_tensor_normalize_and_flip_are_bit_exact(n, h, w, c):
    """ArrayToTensor + Normalize (+ RandomHorizontalFlip): same fp32 operation order -> identical bits"""
    from gdn_pytorch_b200.data import preprocess_u8
    from oracle import transforms as OT
    x = _batch(n, h, w, c, 3)
    flip = (np.arange(n) % 2).astype(np.int32)
    got = preprocess_u8(torch.from_numpy(x).to(dev), torch.from_numpy(flip).to(dev)).cpu()
    want = torch.stack([OT.to_tensor_normalize(OT.flip_scale_crop(x[i], flip[i], None)) for i in range(n)])
    assert got.shape == want.shape and got.dtype == torch.float32
    assert torch.equal(got, want)
    got2 = preprocess_u8(torch.from_numpy(x).to(dev)).cpu()
    want2 = torch.stack([OT.to_tensor_normalize(x[i]) for i in range(n)])
    assert torch.equal(got2, want2)
    assert float(got2.min()) >= -1.0 and float(got2.max()) <= 1.0


@pytest.mark.parametrize("c,lo,hi", [(3, 0, 256), (3, 17, 201), (0, 40, 130), (1, 0, 256), (0, 7, 8)])
def test_scale_crop_is_the_reference_imresize_bit_for_bit(c, lo, hi):
    """RandomScaleCrop (transform_list.py:189-203) on the loader's float32 images: scipy.misc.imresize = per-image min-max
    byte scaling + PIL's 8-bit BILINEAR resize, then the crop -- identical bytes, hence identical fp32 tensors.  Images
    whose range is not exactly [0, 255] (depth PNGs) exercise the stretch; a constant image the max == min case."""
    from gdn_pytorch_b200.data import preprocess_u8, DeviceInputPipeline
    from oracle import transforms as OT
    n, h, w = 6, 64, 96
    rs = np.random.RandomState(5 + c + lo)
    x = rs.randint(lo, hi, size=(n, h, w, c) if c else (n, h, w)).astype(np.uint8)
    pipe = DeviceInputPipeline(dev, train=True, seed=1)
    flip, crop = pipe.draw(n, h, w)
    crop[0] = (h, w, 0, 0)                       # both passes skipped
    crop[1] = (h, crop[1][1], 0, crop[1][3])     # horizontal pass only
    crop[2] = (crop[2][0], w, crop[2][2], 0)     # vertical pass only
    assert ((crop[:, 0] >= h) & (crop[:, 0] <= int(h * 1.15)) & (crop[:, 2] >= 0) & (crop[:, 2] <= crop[:, 0] - h)).all()
    got = preprocess_u8(torch.from_numpy(x).to(dev), torch.from_numpy(flip).to(dev), torch.from_numpy(crop).to(dev)).cpu()
    want = torch.stack([OT.to_tensor_normalize(OT.flip_scale_crop(x[i], flip[i], crop[i])) for i in range(n)])
    assert got.shape == want.shape
    assert torch.equal(got, want)
    if hi - lo > 1 and (lo, hi) != (0, 256):     # the stretch is visible: the zoomed image spans the full byte range
        assert float(got.max()) > 0.9 and float(got.min()) < -0.9


def test_pipeline_yields_what_the_training_step_consumes():
    from gdn_pytorch_b200.data import DeviceInputPipeline
    n, h, w = 2, 32, 64
    gt, rgb, sp = _batch(n, h, w, 0, 1), _batch(n, h, w, 3, 2), _batch(n, h, w, 0, 3)
    for train in (False, True):
        pipe = DeviceInputPipeline(dev, train=train, seed=0)
        g, r, s = pipe(torch.from_numpy(gt).pin_memory(), torch.from_numpy(rgb).pin_memory(), torch.from_numpy(sp).pin_memory())
        assert g.shape == (n, 1, h, w) and r.shape == (n, 3, h, w) and s.shape == (n, 1, h, w)
        assert r.is_cuda and r.dtype == torch.float32
    # the same random draw is applied to all three tensors of a sample (datasets_list.py:88-92)
    pipe = DeviceInputPipeline(dev, train=True, seed=4)
    a, b, _ = pipe(torch.from_numpy(rgb[..., 0].copy()), torch.from_numpy(rgb), None)
    assert torch.equal(a[:, 0], b[:, 0])


def test_host_batch_prefetcher_delivers_batches_in_order():
    """double-buffered H2D feed: every batch arrives intact and in order while 'steps' run on the consumer stream"""
    from gdn_pytorch_b200.data import HostBatchPrefetcher
    feed = HostBatchPrefetcher("cuda")
    batches = [(torch.full((4, 3, 32, 64), float(i)).pin_memory(), torch.full((4, 1, 32, 64), -float(i)).pin_memory(), None)
               for i in range(7)]
    feed.submit(*batches[0])
    acc = torch.zeros((), device="cuda", dtype=torch.float64)
    for i in range(7):
        a, b, c = feed.next()
        if i + 1 < 7:
            feed.submit(*batches[i + 1])
        assert c is None
        torch.cuda._sleep(2_000_000)                       # a "step" that is still running when the next copy starts
        acc += a.double().mean() * 10 + b.double().mean()
        assert float(a[0, 0, 0, 0]) == float(i) and float(b[-1, 0, -1, -1]) == -float(i)
    assert abs(float(acc) - sum(10 * i - i for i in range(7))) < 1e-9
    with pytest.raises(RuntimeError):
        feed.next()
