"""Restatement of calculate_error.compute_errors (TEST INFRASTRUCTURE).

Reference: /root/reference/src/calculate_error.py:10-103.  fp32 torch ops in the reference's operation
order, so the delta-threshold COUNTS are bit-exact with the reference on identical arrays; besides the
reference's 8 batch-averaged floats this oracle also returns the raw per-image integer counts.
"""
import torch
from .losses import godard_crop


def eigen_metrics(gt_np, gt, pred, crop=True):
    """-> (list of 8 floats [abs_diff, abs_rel, sq_rel, a1, a2, a3, rmse, rmse_log], counts[B,4] int64)
    counts[b] = (n_valid, n<1.25, n<1.25^2, n<1.25^3)."""
    B, _, H, W = gt.shape
    acc = [0.0] * 8
    counts = torch.zeros((B, 4), dtype=torch.int64)
    cm = torch.zeros((H, W), dtype=torch.bool)
    if crop:
        y1, y2, x1, x2 = godard_crop(H, W)
        cm[y1:y2, x1:x2] = True
    for b in range(B):
        g, gn, p = gt[b, 0].float(), gt_np[b, 0].float(), pred[b, 0].float()
        p = (p - p.min()) / (p.max() - p.min())          # :38
        g = (g - g.min()) / (g.max() - g.min())          # :39
        gn = (gn + 1.0) / 2.0                            # :41
        g, p, gn = g * 80, p * 80, gn * 80               # :44-46
        valid = (gn < 80) & (g < 80) & (gn > 1) & (g > 1)  # :78
        if crop:
            valid = valid & cm
        vg, vp = g[valid], p[valid]
        vp = vp * torch.median(vg) / torch.median(vp)    # :86 (lower median, left-to-right)
        vp = vp.clamp(1, 80)
        thr = torch.max(vg / vp, vp / vg)
        n = vg.numel()
        c1, c2, c3 = int((thr < 1.25).sum()), int((thr < 1.25 ** 2).sum()), int((thr < 1.25 ** 3).sum())
        counts[b] = torch.tensor([n, c1, c2, c3])
        d = vg - vp
        acc[0] += d.abs().mean().item()
        acc[1] += (d.abs() / vg).mean().item()
        acc[2] += ((d ** 2) / vg).mean().item()
        acc[3] += (thr < 1.25).float().mean().item()
        acc[4] += (thr < 1.25 ** 2).float().mean().item()
        acc[5] += (thr < 1.25 ** 3).float().mean().item()
        acc[6] += torch.sqrt((d ** 2).mean()).item()
        acc[7] += torch.sqrt(((torch.log(vg) - torch.log(vp)) ** 2).mean()).item()
    return [a / B for a in acc], counts


def nyu_metrics(gt, pred, crop=True):
    """Restatement of calculate_error.compute_errors_NYU (/root/reference/src/calculate_error.py:105-151).
    -> (list of 8 floats [abs_diff, abs_rel, log10, a1, a2, a3, rmse, rmse_log], counts[B,4] int64)"""
    B, _, H, W = gt.shape
    acc = [0.0] * 8
    counts = torch.zeros((B, 4), dtype=torch.int64)
    cm = torch.zeros((H, W), dtype=torch.bool)
    if crop:
        y1, y2 = int(0.0359477 * H), int(0.96405229 * H)          # :111
        x1, x2 = int(0.0359477 * W), int(0.96405229 * W)          # :112
        cm[y1:y2, x1:x2] = True
    for b in range(B):
        g, p = gt[b, 0].float(), pred[b, 0].float()
        p = (p - p.min()) / (p.max() - p.min())                    # :121
        g = (g - g.min()) / (g.max() - g.min())                    # :122
        g, p = g * 10, p * 10                                      # :125-126
        valid = (g < 10) & (g > 0)                                 # :128
        if crop:
            valid = valid & cm
        vg, vp = g[valid], p[valid]
        vp = vp * torch.median(vg) / torch.median(vp)              # :134
        vp = vp.clamp(1e-3, 10)                                    # :135
        thr = torch.max(vg / vp, vp / vg)
        n = vg.numel()
        counts[b] = torch.tensor([n, int((thr < 1.25).sum()), int((thr < 1.25 ** 2).sum()), int((thr < 1.25 ** 3).sum())])
        d = vg - vp
        acc[0] += d.abs().mean().item()
        acc[1] += (d.abs() / vg).mean().item()
        acc[2] += (torch.log10(vg) - torch.log10(vp)).abs().mean().item()   # :149
        acc[3] += (thr < 1.25).float().mean().item()
        acc[4] += (thr < 1.25 ** 2).float().mean().item()
        acc[5] += (thr < 1.25 ** 3).float().mean().item()
        acc[6] += torch.sqrt((d ** 2).mean()).item()
        acc[7] += torch.sqrt(((torch.log(vg) - torch.log(vp)) ** 2).mean()).item()
    return [a / B for a in acc], counts


def make3d_metrics(gt_np, gt, pred):
    """Restatement of calculate_error.compute_errors_Make3D (/root/reference/src/calculate_error.py:153-182).
    -> (list of 4 floats [abs_diff, abs_rel, ave_log10, rmse], n_valid[B] int64)"""
    B = gt.shape[0]
    acc = [0.0] * 4
    nv = torch.zeros(B, dtype=torch.int64)
    for b in range(B):
        g, gn, p = gt[b, 0].float(), gt_np[b, 0].float(), pred[b, 0].float()
        g = (g - g.min()) / (g.max() - g.min())                    # :163
        p = (p - p.min()) / (p.max() - p.min())                    # :164
        gn = (gn - gn.min()) / (gn.max() - gn.min())               # :165
        valid = (gn > 1e-2) & (g > 1e-2)                           # :166
        g, p, gn = g * 80, p * 80, gn * 80                         # :167-169
        valid = valid & (gn < 80) & (g < 80)                       # :171
        vg = g[valid].clamp(1e-2, 80)                              # :173
        vp = p[valid].clamp(1e-2, 80)                              # :174
        vp = vp * torch.median(vg) / torch.median(vp)              # :175
        nv[b] = vg.numel()
        d = vg - vp
        acc[3] += torch.sqrt((d ** 2).mean()).item()
        acc[2] += (torch.log10(vg) - torch.log10(vp)).abs().mean().item()
        acc[0] += d.abs().mean().item()
        acc[1] += (d.abs() / vg).mean().item()
    return [a / B for a in acc], nv
