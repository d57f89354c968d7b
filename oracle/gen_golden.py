"""Generate tests/golden/*.npz by running the UNMODIFIED reference (dev container only; TEST INFRA).

    python -m oracle.gen_golden

What is recorded (all fp32, CPU, torch threads = default):
  net_<Class>_{eval,train}.npz : reference forward at 2x{1|3}x32x64 with synth_state_dict(seed 0): the depth
                                 output in full, and (mean, abs-mean, 5 samples) of the other 7 returned maps
  init_<Class>.npz             : per-key (sum, abs-sum) of the reference's own random init under manual_seed(0)
  loss.npz                     : utils.imgrad_loss / depth_smoothness and the inline trainer loss terms
                                 (trainer.py is not importable; its arithmetic is replayed line by line here with
                                 the reference's own helpers and boolean-mask writes, see _ref_inline_*)
  metrics.npz                  : calculate_error.compute_errors on synthetic pred/gt
  metrics_variants.npz         : calculate_error.compute_errors_NYU / compute_errors_Make3D on synthetic pred/gt
  trainstep_DtoD.npz           : one fwd+loss+bwd+Adam step of the reference AutoEncoder_DtoD at 2x1x32x64
"""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.refimport import load_reference  # noqa: E402
from oracle import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
H, W, B = 32, 64, 2
IDX = [0, 7, 101, 1009, -1]


def summarize(t):
    f = t.detach().reshape(-1).double()
    return np.array([f.mean().item(), f.abs().mean().item()] + [f[i % f.numel()].item() for i in IDX])


def _ref_inline_berhu(outputs, depths, sparse):
    """trainer.py:705-720 replayed verbatim (boolean-mask in-place writes and all)"""
    Hh, Ww = depths.size(2), depths.size(3)
    y1, y2 = int(0.40810811 * Hh), int(0.99189189 * Hh)      # trainer.py:644
    x1, x2 = int(0.03594771 * Ww), int(0.96405229 * Ww)      # trainer.py:645
    crop_mask = depths != depths
    crop_mask[:, :, y1:y2, x1:x2] = 1
    valid_mask = sparse > -1
    valid_mask = valid_mask[:, 0, :, :].unsqueeze(1)
    diff = outputs - depths
    diff_abs = torch.abs(diff)
    diff_2 = torch.pow(outputs - depths, 2)
    c = 0.2 * torch.max(diff_abs.detach())
    mask2 = torch.gt(diff_abs.detach(), c)
    diff_abs[mask2] = (diff_2[mask2] + (c * c)) / (2 * c)
    diff_abs[~crop_mask] = 0.1 * diff_abs[~crop_mask]
    diff_abs[crop_mask & (~valid_mask)] = 0.3 * diff_abs[crop_mask & (~valid_mask)]
    return 3 * diff_abs.mean(), c


def gen_metric_variants(ce):
    """metrics_variants.npz: calculate_error.compute_errors_NYU / compute_errors_Make3D on synthetic pred / gt"""
    rec = {}
    for hh, ww, tag in ((128, 416, "kitti"), (48, 64, "small")):
        pred = synth.synth_pred(3, hh, ww, 6)
        gt = synth.synth_depth(3, hh, ww, 6)
        gtn = synth.synth_sparse(gt, 6, keep=0.6)
        rec["nyu_" + tag] = np.array(ce.compute_errors_NYU(gt.clone(), pred.clone(), crop=True), dtype=np.float64)
        rec["nyu_nocrop_" + tag] = np.array(ce.compute_errors_NYU(gt.clone(), pred.clone(), crop=False), dtype=np.float64)
        rec["make3d_" + tag] = np.array(ce.compute_errors_Make3D(gtn.clone(), gt.clone(), pred.clone()), dtype=np.float64)
        print("metric variants", tag, rec["nyu_" + tag], rec["make3d_" + tag])
    np.savez(os.path.join(OUT, "metrics_variants.npz"), **rec)


def main():
    os.makedirs(OUT, exist_ok=True)
    ae, ce, ut = load_reference()
    torch.Tensor.cuda = lambda self, *a, **k: self  # AutoEncoder.forward calls x.cuda() (AE_model_unet.py:161)
    torch.manual_seed(0)

    # ---------------------------------------------------------------- networks
    for name, cin in (("AutoEncoder_2", 3), ("AutoEncoder_DtoD", 1), ("AutoEncoder", 3)):
        cls = getattr(ae, name)
        torch.manual_seed(0)
        m = cls(height=H, width=W)
        init = {k: np.array([v.double().sum().item(), v.double().abs().sum().item()])
                for k, v in m.state_dict().items()}
        np.savez(os.path.join(OUT, "init_%s.npz" % name), **init)
        shapes = {k: v.shape for k, v in m.state_dict().items()}
        sd = synth.synth_state_dict(shapes, seed=0)
        x = synth.synth_rgb(B, H, W, 0) if cin == 3 else synth.synth_depth(B, H, W, 0)
        for mode in ("eval", "train"):
            m.load_state_dict({k: v.clone() for k, v in sd.items()})
            m.train(mode == "train")
            with torch.no_grad():
                outs = m(x, istrain=True)
            rec = {"depth": outs[7].numpy()}
            for i in range(7):
                rec["t%d" % i] = summarize(outs[i])
                rec["t%d_shape" % i] = np.array(outs[i].shape)
            if mode == "train":  # running stats after one train-mode forward
                st = m.state_dict()
                for k in ("res64_down1.main.1.running_mean", "res64_down1.main.1.running_var",
                          "res512_3.main.4.running_var"):
                    rec["rs_" + k] = st[k].numpy()
            np.savez(os.path.join(OUT, "net_%s_%s.npz" % (name, mode)), **rec)
            print(name, mode, "depth absmax", float(outs[7].abs().max()))

    # ------------------------------------------------------------------ losses
    out = synth.synth_pred(B, H, W, 3).requires_grad_(True)
    dep = synth.synth_depth(B, H, W, 0)
    spa = synth.synth_sparse(dep, 0)
    rgb = synth.synth_rgb(B, H, W, 0)
    rec = {}
    l = ut.imgrad_loss(out, dep)
    rec["imgrad_loss"] = l.item()
    rec["imgrad_grad"] = torch.autograd.grad(l, out)[0].numpy()
    sm = torch.mean(torch.abs(0.1 * ut.depth_smoothness(out, rgb)))   # trainer.py:753-754
    rec["smooth_loss"] = sm.item()
    rec["smooth_grad"] = torch.autograd.grad(sm, out)[0].numpy()
    bl, c = _ref_inline_berhu(out, dep, spa)
    rec["berhu_loss"] = bl.item()
    rec["berhu_c"] = c.item()
    rec["berhu_grad"] = torch.autograd.grad(bl, out)[0].numpy()
    # full DtoD loss (trainer.py:433-456) and RtoD loss without latent term (:757)
    bl2, _ = _ref_inline_berhu(out, dep, spa)
    dl = bl2 + 3 * ut.imgrad_loss(out, dep.detach())
    rec["dtod_loss"] = dl.item()
    rec["dtod_grad"] = torch.autograd.grad(dl, out)[0].numpy()
    bl3, _ = _ref_inline_berhu(out, dep, spa)
    rl = bl3 + torch.mean(torch.abs(0.1 * ut.depth_smoothness(out, rgb)))
    rec["rtod_nolatent_loss"] = rl.item()
    rec["rtod_nolatent_grad"] = torch.autograd.grad(rl, out)[0].numpy()
    np.savez(os.path.join(OUT, "loss.npz"), **rec)

    # ----------------------------------------------------------------- metrics
    rec = {}
    for hh, ww, tag in ((128, 416, "kitti"), (32, 64, "small")):
        pred = synth.synth_pred(4, hh, ww, 5)
        gt = synth.synth_depth(4, hh, ww, 5)
        gtn = synth.synth_sparse(gt, 5, keep=0.6)
        res = ce.compute_errors(gtn, gt, pred, crop=True)
        rec[tag] = np.array(res, dtype=np.float64)
        print("metrics", tag, res)
    np.savez(os.path.join(OUT, "metrics.npz"), **rec)
    gen_metric_variants(ce)

    # ------------------------------------------------- one DtoD training step
    torch.manual_seed(0)
    m = ae.AutoEncoder_DtoD(height=H, width=W)
    shapes = {k: v.shape for k, v in m.state_dict().items()}
    m.load_state_dict(synth.synth_state_dict(shapes, seed=1))
    m.train()
    opt = torch.optim.Adam(m.parameters(), 2e-5, [0.9, 0.999], eps=1e-8, weight_decay=5e-4)  # GDN_main.py:157
    outputs = m(dep, istrain=False)
    bl, _ = _ref_inline_berhu(outputs, dep, spa)
    loss = bl + 3 * ut.imgrad_loss(outputs, dep.detach())
    opt.zero_grad()
    loss.backward()
    rec = {"loss": loss.item()}
    params = dict(m.named_parameters())
    for k in ("downconv0.main.1.weight", "res64_down1.main.0.weight", "res512_3.main.4.weight",
              "upconv1.main.0.weight", "upconv4.weight"):
        rec["g_" + k] = summarize(params[k].grad)
    opt.step()
    for k in ("downconv0.main.1.weight", "res512_3.main.4.weight", "upconv4.weight"):
        rec["p_" + k] = summarize(params[k].data)
    np.savez(os.path.join(OUT, "trainstep_DtoD.npz"), **rec)
    print("trainstep loss", loss.item())


if __name__ == "__main__":
    main()
