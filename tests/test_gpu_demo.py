"""GPU (-m gpu): the demo path -- /root/reference/src/depth_extract.py:60-147.

bytescale / PIL-exact resize kernels bit-for-bit against oracle/imresize.py (itself pinned on Pillow), and the
whole DepthExtractor: pre-processing exact, network within 1e-2 of the fp32 oracle, post-processing exact given the
network's own output, CUDA-graph replay identical to the eager pass."""
import numpy as np
import pytest
import torch

from tests.test_gpu_network import _module, dev

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("h,w,c,oh,ow", [(375, 1242, 3, 128, 416), (128, 416, 1, 375, 1242), (37, 53, 3, 128, 416),
                                          (200, 300, 3, 100, 300), (100, 50, 1, 33, 77), (480, 640, 3, 128, 416),
                                          (128, 416, 1, 128, 416), (9, 7, 4, 64, 3)])
def test_resize_u8_bit_exact(h, w, c, oh, ow):
    from gdn_pytorch_b200.demo import resize_u8
    from oracle import imresize as OI
    img = np.random.RandomState(h + w).randint(0, 256, (h, w, c)).astype(np.uint8)
    a = img[:, :, 0] if c == 1 else img
    got = resize_u8(torch.from_numpy(a).to(dev), (oh, ow)).cpu().numpy()
    assert np.array_equal(got, OI.resize_u8(a, (oh, ow)))


def test_resize_u8_batch_and_golden():
    """(N, H, W, C) batches resize image by image; case 0 of the Pillow fixtures through the device kernels"""
    from gdn_pytorch_b200.demo import resize_u8
    from oracle.gen_golden_imresize import CASES, case_input, digest
    from tests.util import golden
    gold = golden("imresize.npz")
    for i, (seed, h, w, c, oh, ow) in enumerate(CASES):
        got = resize_u8(torch.from_numpy(case_input(seed, h, w, c)).to(dev), (oh, ow)).cpu().numpy()
        assert digest(got) == str(gold["case%d_sha256" % i]), i
    rng = np.random.RandomState(5)
    batch = rng.randint(0, 256, (3, 40, 60, 3)).astype(np.uint8)
    got = resize_u8(torch.from_numpy(batch).to(dev), (17, 90)).cpu().numpy()
    from oracle import imresize as OI
    for n in range(3):
        assert np.array_equal(got[n], OI.resize_u8(batch[n], (17, 90)))


@pytest.mark.parametrize("f64", [False, True])
def test_bytescale_bit_exact(f64):
    from gdn_pytorch_b200.demo import bytescale
    from oracle import imresize as OI
    rng = np.random.RandomState(11)
    f = np.tanh(rng.randn(128, 416)).astype(np.float32)
    got = bytescale(torch.from_numpy(f).to(dev), f64=f64).cpu().numpy()
    ref = OI.bytescale(f.astype(np.float64) if f64 else f)
    assert np.array_equal(got, ref)
    u = rng.randint(7, 201, (50, 70, 3)).astype(np.uint8)          # uint8 storage of a float image (load_as_float)
    got = bytescale(torch.from_numpy(u).to(dev), f64=f64).cpu().numpy()
    assert np.array_equal(got, OI.bytescale(u.astype(np.float64 if f64 else np.float32)))
    const = torch.full((8, 8), 0.25, device=dev)
    assert int(bytescale(const, f64=f64).max()) == 0


@pytest.mark.parametrize("name", ["AutoEncoder", "AutoEncoder_2"])
def test_depth_extractor_matches_the_reference_pipeline(name):
    from gdn_pytorch_b200.demo import DepthExtractor
    from oracle import imresize as OI, model as OM
    m, sd = _module(name, seed=3, h=128, w=416)
    m.eval()
    ex = DepthExtractor(m)
    rng = np.random.RandomState(2)
    outs = []
    for i, (oh, ow) in enumerate([(375, 1242), (370, 1226), (375, 1242), (240, 320)]):   # eager, capture, replays
        low = rng.randint(0, 256, (oh // 8 + 1, ow // 8 + 1, 3)).astype(np.uint8)
        img = OI.resize_u8(low, (oh, ow))                                       # a smooth-ish synthetic photograph
        out = ex(img)
        assert out.shape == (oh, ow) and out.dtype == torch.uint8 and out.is_cuda
        x = torch.from_numpy(OI.demo_preprocess(img))
        assert torch.equal(ex.static_in.cpu(), x), "pre-processing differs from imresize + normalise"
        depth = ex.last_depth.detach().cpu()
        with torch.no_grad():
            ref = OM.FORWARDS[name](sd, x, istrain=False)
        err = (depth - ref).abs().max().item() / ref.abs().max().item()
        assert err <= 1e-2, (i, err)
        post = OI.demo_postprocess(depth.reshape(128, 416).numpy(), (oh, ow))
        assert np.array_equal(out.cpu().numpy(), post), "post-processing differs from imresize of the same depth map"
        outs.append((img, depth.clone()))
    assert ex._graph is not None
    # the replayed graph computes what the eager plan computes
    eager = DepthExtractor(m, use_graph=False)
    eager(outs[-1][0])
    assert torch.equal(eager.last_depth.cpu(), outs[-1][1])


def test_depth_extractor_refuses_cpu_and_train_mode():
    from gdn_pytorch_b200.demo import DepthExtractor, bytescale
    m, _ = _module("AutoEncoder", seed=3, h=128, w=416)
    m.train()
    with pytest.raises(RuntimeError):
        DepthExtractor(m)
    with pytest.raises(RuntimeError):
        bytescale(torch.zeros(4, 4))
