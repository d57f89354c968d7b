"""CPU: pin the oracle (oracle/*.py) against the golden vectors produced by the UNMODIFIED reference
(oracle/gen_golden.py; fixtures in tests/golden/), and live against /root/reference when it is present."""
import numpy as np
import pytest
import torch

from oracle import model as OM, losses as OL, metrics as OMet, adam as OA, synth
from oracle.refimport import reference_available
from tests.util import golden, shapes_of, build_module

H, W, B = 32, 64, 2
IDX = [0, 7, 101, 1009, -1]


def summarize(t):
    f = t.detach().reshape(-1).double()
    return np.array([f.mean().item(), f.abs().mean().item()] + [f[i % f.numel()].item() for i in IDX])


@pytest.mark.parametrize("name,cin", [("AutoEncoder_2", 3), ("AutoEncoder_DtoD", 1), ("AutoEncoder", 3)])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_network_forward_matches_reference_golden(name, cin, mode):
    gold = golden("net_%s_%s.npz" % (name, mode))
    sd = synth.synth_state_dict(shapes_of(name), seed=0)
    x = synth.synth_rgb(B, H, W, 0) if cin == 3 else synth.synth_depth(B, H, W, 0)
    with torch.no_grad():
        outs = OM.FORWARDS[name](sd, x, istrain=True, train=(mode == "train"), update_running=(mode == "train"))
    tol = 2e-5 if mode == "eval" else 2e-3   # train mode: fp32 noise is amplified by batch-stat BN at random init
    d = torch.from_numpy(gold["depth"])
    assert (outs[7] - d).abs().max().item() <= tol * max(1.0, d.abs().max().item())
    for i in range(7):
        assert tuple(outs[i].shape) == tuple(gold["t%d_shape" % i])
        np.testing.assert_allclose(summarize(outs[i])[:2], gold["t%d" % i][:2], rtol=50 * tol, atol=1e-6)
    if mode == "train":
        for k in ("res64_down1.main.1.running_mean", "res64_down1.main.1.running_var", "res512_3.main.4.running_var"):
            np.testing.assert_allclose(sd[k].numpy(), gold["rs_" + k], rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("name", ["AutoEncoder_2", "AutoEncoder_DtoD", "AutoEncoder"])
def test_default_init_matches_reference(name):
    """product modules constructed under manual_seed(0) reproduce the reference's random init bit for bit"""
    gold = golden("init_%s.npz" % name)
    torch.manual_seed(0)
    m = build_module(name, height=H, width=W)
    sd = m.state_dict()
    assert sorted(sd.keys()) == sorted(gold.files)
    for k, v in sd.items():
        got = np.array([v.double().sum().item(), v.double().abs().sum().item()])
        np.testing.assert_allclose(got, gold[k], rtol=1e-12, atol=0)


def _loss_inputs():
    out = synth.synth_pred(B, H, W, 3).requires_grad_(True)
    dep = synth.synth_depth(B, H, W, 0)
    spa = synth.synth_sparse(dep, 0)
    rgb = synth.synth_rgb(B, H, W, 0)
    return out, dep, spa, rgb


def test_loss_terms_match_reference_golden():
    gold = golden("loss.npz")
    out, dep, spa, rgb = _loss_inputs()
    l = OL.imgrad_loss(out, dep)
    assert abs(l.item() - gold["imgrad_loss"]) < 1e-6 * abs(gold["imgrad_loss"])
    np.testing.assert_allclose(torch.autograd.grad(l, out)[0].numpy(), gold["imgrad_grad"], atol=1e-9)
    sm = (0.1 * OL.depth_smoothness(out, rgb)).abs().mean()
    assert abs(sm.item() - gold["smooth_loss"]) < 1e-6 * abs(gold["smooth_loss"])
    np.testing.assert_allclose(torch.autograd.grad(sm, out)[0].numpy(), gold["smooth_grad"], atol=1e-9)
    bl, c = OL.berhu_masked(out, dep, spa)
    assert abs(bl.item() - gold["berhu_loss"]) < 1e-6 * abs(gold["berhu_loss"])
    assert abs(c.item() - gold["berhu_c"]) < 1e-7
    np.testing.assert_allclose(torch.autograd.grad(bl, out)[0].numpy(), gold["berhu_grad"], atol=1e-9)
    d = OL.dtod_loss(out, dep, spa)
    assert abs(d["loss"].item() - gold["dtod_loss"]) < 1e-6 * abs(gold["dtod_loss"])
    np.testing.assert_allclose(torch.autograd.grad(d["loss"], out)[0].numpy(), gold["dtod_grad"], atol=1e-9)
    r = OL.rtod_loss(out, dep, spa, rgb)
    assert abs(r["loss"].item() - gold["rtod_nolatent_loss"]) < 1e-6 * abs(gold["rtod_nolatent_loss"])
    np.testing.assert_allclose(torch.autograd.grad(r["loss"], out)[0].numpy(), gold["rtod_nolatent_grad"], atol=1e-9)


def test_metrics_match_reference_golden():
    gold = golden("metrics.npz")
    for hh, ww, tag in ((128, 416, "kitti"), (32, 64, "small")):
        pred = synth.synth_pred(4, hh, ww, 5)
        gt = synth.synth_depth(4, hh, ww, 5)
        gtn = synth.synth_sparse(gt, 5, keep=0.6)
        res, counts = OMet.eigen_metrics(gtn, gt, pred, crop=True)
        np.testing.assert_allclose(np.array(res), gold[tag], rtol=2e-6)
        # a1..a3 are count/n in fp32: identical to the reference to the last bit
        for j in range(3):
            a = sum(np.float32(counts[b, 1 + j].item()) / np.float32(counts[b, 0].item()) for b in range(4)) / 4
            assert abs(a - gold[tag][3 + j]) < 1e-7


def test_dtod_train_step_matches_reference_golden():
    """fwd + loss + bwd (autograd through the functional oracle) + restated Adam vs the reference's own step"""
    gold = golden("trainstep_DtoD.npz")
    name = "AutoEncoder_DtoD"
    sd = synth.synth_state_dict(shapes_of(name), seed=1)
    m = build_module(name, init_weights=False, height=H, width=W)
    pnames = [n for n, _ in m.named_parameters()]
    for k in pnames:
        sd[k].requires_grad_(True)
    dep = synth.synth_depth(B, H, W, 0)
    spa = synth.synth_sparse(dep, 0)
    out = OM.autoencoder_dtod(sd, dep, istrain=False, train=True)
    terms = OL.dtod_loss(out, dep, spa)
    assert abs(terms["loss"].item() - gold["loss"]) < 2e-4 * abs(gold["loss"])
    grads = torch.autograd.grad(terms["loss"], [sd[k] for k in pnames])
    gmap = dict(zip(pnames, grads))
    for k in ("downconv0.main.1.weight", "res64_down1.main.0.weight", "res512_3.main.4.weight",
              "upconv1.main.0.weight", "upconv4.weight"):
        np.testing.assert_allclose(summarize(gmap[k])[1], gold["g_" + k][1], rtol=5e-3)
    for k in ("downconv0.main.1.weight", "res512_3.main.4.weight", "upconv4.weight"):
        p = sd[k].detach().clone()
        OA.adam_step(p, gmap[k], torch.zeros_like(p), torch.zeros_like(p), 1, 2e-5)
        np.testing.assert_allclose(summarize(p), gold["p_" + k], rtol=1e-4, atol=1e-7)


@pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")
def test_oracle_live_against_reference():
    """dev container only: run the imported reference and the oracle side by side at 32x64"""
    from oracle.refimport import load_reference
    import contextlib, io
    ae, ce, ut = load_reference()
    name = "AutoEncoder_2"
    with contextlib.redirect_stdout(io.StringIO()):
        ref = ae.AutoEncoder_2(height=H, width=W)
    sd = synth.synth_state_dict(shapes_of(name), seed=3)
    ref.load_state_dict({k: v.clone() for k, v in sd.items()})
    ref.eval()
    x = synth.synth_rgb(B, H, W, 9)
    with torch.no_grad():
        a = ref(x, istrain=False)
        b = OM.autoencoder_2(sd, x)
    assert (a - b).abs().max().item() < 1e-5
    pred, gt = synth.synth_pred(3, 128, 416, 8), synth.synth_depth(3, 128, 416, 8)
    gtn = synth.synth_sparse(gt, 8, keep=0.5)
    r1 = ce.compute_errors(gtn, gt, pred, crop=True)
    r2, _ = OMet.eigen_metrics(gtn, gt, pred, crop=True)
    np.testing.assert_allclose(np.array(r1), np.array(r2), rtol=2e-6)


def test_metric_variants_match_reference_golden():
    """compute_errors_NYU / compute_errors_Make3D restatements against the reference's own outputs"""
    gold = golden("metrics_variants.npz")
    for hh, ww, tag in ((128, 416, "kitti"), (48, 64, "small")):
        pred = synth.synth_pred(3, hh, ww, 6)
        gt = synth.synth_depth(3, hh, ww, 6)
        gtn = synth.synth_sparse(gt, 6, keep=0.6)
        np.testing.assert_allclose(np.array(OMet.nyu_metrics(gt, pred, True)[0]), gold["nyu_" + tag], rtol=5e-6)
        np.testing.assert_allclose(np.array(OMet.nyu_metrics(gt, pred, False)[0]), gold["nyu_nocrop_" + tag], rtol=5e-6)
        np.testing.assert_allclose(np.array(OMet.make3d_metrics(gtn, gt, pred)[0]), gold["make3d_" + tag], rtol=5e-6)


@pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")
def test_transform_restatement_matches_reference_classes():
    """oracle/transforms.py against the reference's own ArrayToTensor / Normalize / RandomHorizontalFlip classes
    (transform_list.py imported with scipy.misc stubbed: imresize no longer exists, so RandomScaleCrop cannot run)"""
    import importlib.util
    import random
    import sys
    import types
    from oracle import transforms as OT
    if "scipy.misc" not in sys.modules or not hasattr(sys.modules["scipy.misc"], "imresize"):
        m = types.ModuleType("scipy.misc")
        m.imresize = m.imrotate = None
        sys.modules["scipy.misc"] = m
    spec = importlib.util.spec_from_file_location("_gdn_reference_transform_list", "/root/reference/src/transform_list.py")
    tl = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tl)
    rs = np.random.RandomState(0)
    h, w = 32, 64
    gt = rs.randint(0, 256, (h, w)).astype(np.float32)          # load_as_float yields float arrays of 0..255
    rgb = rs.randint(0, 256, (h, w, 3)).astype(np.float32)
    pipe = tl.Compose([tl.ArrayToTensor(height=h, width=w), tl.Normalize(mean=[0.5, 0.5, 0.5], std=[0.5, 0.5, 0.5])])
    ref = pipe([gt.copy(), rgb.copy()])
    assert torch.equal(ref[0], OT.to_tensor_normalize(gt.astype(np.uint8)))
    assert torch.equal(ref[1], OT.to_tensor_normalize(rgb.astype(np.uint8)))
    random.seed(1)                                               # first draw 0.134 < 0.5 -> flips
    flipped = tl.RandomHorizontalFlip()([rgb.copy()])[0]
    assert np.array_equal(flipped.astype(np.uint8), OT.flip_scale_crop(rgb.astype(np.uint8), True, None))


# ---------------------------------------------------------------------------------------- demo path: imresize
def test_imresize_oracle_matches_pillow_golden():
    """oracle/imresize.py::resize_u8 == PIL.Image.resize(BILINEAR), bit-exact, on the committed fixtures
    (tests/golden/imresize.npz, written by oracle/gen_golden_imresize.py from Pillow itself)"""
    from oracle import imresize as OI
    from oracle.gen_golden_imresize import CASES, case_input, digest
    gold = golden("imresize.npz")
    for i, (seed, h, w, c, oh, ow) in enumerate(CASES):
        got = OI.resize_u8(case_input(seed, h, w, c), (oh, ow))
        assert np.array_equal(got[:24, :24], gold["case%d_corner" % i]), i
        assert digest(got) == str(gold["case%d_sha256" % i]), i


def test_imresize_oracle_matches_pillow_live():
    """same, against the Pillow in this environment when it imports (other sizes than the fixtures)"""
    Image = pytest.importorskip("PIL.Image")
    from oracle import imresize as OI
    rng = np.random.RandomState(7)
    for (h, w, c, oh, ow) in [(64, 96, 3, 23, 200), (31, 17, 1, 31, 40), (90, 120, 3, 128, 416), (16, 16, 1, 5, 3)]:
        img = rng.randint(0, 256, (h, w, c)).astype(np.uint8)
        img = img[:, :, 0] if c == 1 else img
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
        assert np.array_equal(OI.resize_u8(img, (oh, ow)), ref), (h, w, c, oh, ow)


def test_bytescale_oracle_properties():
    """scipy.misc.bytescale restatement (unpinned: SciPy >= 1.3 has no imresize): uint8 passes through, the extremes
    map to 0 / 255, a constant image maps to 0, float64 and float32 evaluation agree except at .5 boundaries"""
    from oracle import imresize as OI
    rng = np.random.RandomState(3)
    a = rng.randint(0, 256, (20, 30)).astype(np.uint8)
    assert OI.bytescale(a) is a or np.array_equal(OI.bytescale(a), a)
    f = np.tanh(rng.randn(40, 50)).astype(np.float32)
    b = OI.bytescale(f)
    assert b.dtype == np.uint8 and b.min() == 0 and b.max() == 255
    assert OI.bytescale(np.full((4, 4), 3.5, np.float32)).max() == 0
    full = np.arange(256, dtype=np.float32).reshape(16, 16)
    assert np.array_equal(OI.bytescale(full), np.arange(256).reshape(16, 16))       # identity on a full-range image
    d = np.abs(OI.bytescale(f.astype(np.float64)).astype(int) - b.astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.01
