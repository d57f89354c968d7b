"""fp32 restatement of the reference's inline training losses (TEST INFRASTRUCTURE).

Reference: /root/reference/src/trainer.py:433-456 (DtoD), :644-656 + :705-757 (RtoD),
           /root/reference/src/utils.py:105-131 (imgrad, imgrad_loss), :139-178 (gradient_x/y, depth_smoothness).
All functions are differentiable w.r.t. ``outputs`` through torch autograd.
"""
import torch
import torch.nn.functional as F


def garg_crop(H, W):
    """rows/cols of the Garg ECCV16 crop used by the TRAINING loss (trainer.py:644-645)"""
    return int(0.40810811 * H), int(0.99189189 * H), int(0.03594771 * W), int(0.96405229 * W)


def godard_crop(H, W):
    """crop used by compute_errors (calculate_error.py:28-29)"""
    return int(0.3324324 * H), int(0.91351351 * H), int(0.0359477 * W), int(0.96405229 * W)


def berhu_masked(outputs, depths, sparse):
    """masked BerHu term, trainer.py:705-720 (RtoD) == :433-448 (DtoD).
    Returns (output_loss, c).  ``sparse`` may be None (non-KITTI: no crop / validity weights)."""
    diff = outputs - depths
    a = diff.abs()
    c = 0.2 * a.detach().max()
    sq = (diff * diff + c * c) / (2 * c)
    val = torch.where(a.detach() > c, sq, a)
    if sparse is not None:
        B, _, H, W = outputs.shape
        y1, y2, x1, x2 = garg_crop(H, W)
        crop = torch.zeros((B, 1, H, W), dtype=torch.bool, device=outputs.device)
        crop[:, :, y1:y2, x1:x2] = True
        valid = sparse[:, 0:1] > -1
        wgt = torch.where(crop, torch.where(valid, 1.0, 0.3), 0.1).to(val.dtype)
        val = val * wgt
    return 3 * val.mean(), c


def sobel_xy(img):
    """utils.imgrad (utils.py:105-127): channel mean, then 3x3 Sobel cross-correlation with zero padding.
    Returns (grad_y, grad_x)."""
    m = img.mean(1, keepdim=True)
    fx = torch.tensor([[1., 0., -1.], [2., 0., -2.], [1., 0., -1.]], device=img.device).view(1, 1, 3, 3)
    fy = torch.tensor([[1., 2., 1.], [0., 0., 0.], [-1., -2., -1.]], device=img.device).view(1, 1, 3, 3)
    return F.conv2d(m, fy, padding=1), F.conv2d(m, fx, padding=1)


def imgrad_loss(pred, gt):
    """utils.imgrad_loss (utils.py:129-133)"""
    gy, gx = sobel_xy(pred)
    ty, tx = sobel_xy(gt)
    return (gy - ty).abs().mean() + (gx - tx).abs().mean()


def depth_smoothness(depth, img):
    """utils.depth_smoothness (utils.py:165-178) with replicate-padded forward differences (:139-149)"""

    def gx(t):
        t = F.pad(t, (0, 1, 0, 0), mode="replicate")
        return t[..., :-1] - t[..., 1:]

    def gy(t):
        t = F.pad(t, (0, 0, 0, 1), mode="replicate")
        return t[..., :-1, :] - t[..., 1:, :]

    wx = torch.exp(-gx(img).abs().mean(1, keepdim=True))
    wy = torch.exp(-gy(img).abs().mean(1, keepdim=True))
    return (gx(depth) * wx).abs() + (gy(depth) * wy).abs()


def dtod_loss(outputs, depths, sparse):
    """DtoD step loss, trainer.py:433-456: 3*mean(BerHu) + 3*imgrad_loss.  Returns dict of terms."""
    out_loss, c = berhu_masked(outputs, depths, sparse)
    grad_loss = 3 * imgrad_loss(outputs, depths.detach())
    return {"loss": out_loss + grad_loss, "output_loss": out_loss, "gradient_loss": grad_loss, "c": c}


LATENT_W = (1.0, 2.5, 14.0, 12.0)


def latent_loss(feats, feats_tar, with_grad=False):
    """trainer.py:726-733; both feature sets come from no_grad DtoD passes (:699-703) -> constant.
    with_grad=True keeps the graph of the prediction features: the paper-faithful opt-in of SURVEY.md 8f row 3
    (what the published code would compute without the no_grad around :702-703); targets stay detached."""
    tot = 0.0
    for w, f, t in zip(LATENT_W, feats, feats_tar):
        tot = tot + w * F.mse_loss(f if with_grad else f.detach(), t.detach())
    return 1.5 * (tot / 4)


def rtod_loss(outputs, depths, sparse, rgb, feats=None, feats_tar=None, guidance_grad=False):
    """RtoD step loss, trainer.py:705-757.  feats/feats_tar: 4 DtoD feature maps each (or None = RtoD_single)."""
    out_loss, c = berhu_masked(outputs, depths, sparse)
    rmse = torch.sqrt(((outputs - depths) ** 2).detach().mean())  # diagnostic, trainer.py:722-723
    lat = latent_loss(feats, feats_tar, guidance_grad) if feats is not None else torch.zeros((), device=outputs.device)
    smooth = (0.1 * depth_smoothness(outputs, rgb)).abs().mean()
    return {"loss": out_loss + lat + smooth, "output_loss": out_loss, "latent_loss": lat,
            "smooth_loss": smooth, "rmse_loss": rmse, "c": c}
