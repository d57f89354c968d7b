"""validate() of the reference (/root/reference/src/trainer.py:17-87) without per-batch host synchronisation.

The reference moves a batch to the GPU, runs ``model(x, istrain=False)`` under no_grad, calls ``compute_errors`` (a
Python loop over images with ~40 launches, two sorts and eight ``.item()`` syncs each) and feeds the eight floats to an
AverageMeter.  Here a batch is one network launch sequence + ONE metric kernel; the eight per-batch numbers stay in a
device buffer and the host reads everything once, in ``result()``.  With several ranks the per-batch rows are
all-gathered at the end ("fused error-metric reduction").
"""
import torch
import torch.distributed as dist

from . import ops

ERROR_NAMES = {
    "KITTI": ['abs_diff', 'abs_rel', 'sq_rel', 'a1', 'a2', 'a3', 'rmse', 'rmse_log'],     # trainer.py:20
    "NYU": ['abs_diff', 'abs_rel', 'log10', 'a1', 'a2', 'a3', 'rmse', 'rmse_log'],        # trainer.py:203
    "Make3D": ['abs_diff', 'abs_rel', 'ave_log10', 'rmse'],
}
_VARIANT = {"KITTI": 0, "NYU": 1, "Make3D": 2}


class DeviceValidator:
    def __init__(self, model, mode="RtoD", dataset="KITTI", crop=True, max_batches=4096, group=None):
        if dataset not in _VARIANT:
            raise ValueError("dataset must be one of %s" % list(_VARIANT))
        self.model, self.mode, self.dataset, self.crop, self.group = model, mode, dataset, crop, group
        self.dev = next(model.parameters()).device
        self.rows = torch.zeros((max_batches, 8), dtype=torch.float64, device=self.dev)
        self.n = 0

    def reset(self):
        self.n = 0

    @torch.no_grad()
    def update(self, depth, img, depth_np):
        """one iteration of the loop at trainer.py:32-45; all arguments fp32 CUDA tensors (what .cuda() yields there)"""
        x = img if self.mode in ("RtoD", "RtoD_test") else depth          # trainer.py:38-41
        out = self.model(x, istrain=False)
        out8, _ = ops._depth_metrics(_VARIANT[self.dataset], None if self.dataset == "NYU" else depth_np, depth, out,
                                     self.crop)
        if self.n >= self.rows.shape[0]:
            raise RuntimeError("DeviceValidator: more than max_batches batches")
        self.rows[self.n].copy_(out8)
        self.n += 1
        return out

    def result(self):
        """-> (errors.avg, min_errors.avg, error_names) like validate() (trainer.py:87).  One device->host read."""
        rows = self.rows[: self.n]
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            parts = [torch.zeros_like(rows) for _ in range(dist.get_world_size(self.group))]
            dist.all_gather(parts, rows.contiguous(), group=self.group)      # equal batch counts per rank
            rows = torch.cat(parts, 0)
        host = rows.cpu().tolist()
        names = ERROR_NAMES[self.dataset]
        cols = [0, 1, 2, 6] if self.dataset == "Make3D" else list(range(8))
        per_batch = [[r[c] for c in cols] for r in host]
        nb = max(len(per_batch), 1)
        avg = [sum(r[i] for r in per_batch) / nb for i in range(len(cols))]
        # "min_errors": the same per-batch values summed in order of increasing abs_diff (trainer.py:62-85)
        order = sorted(range(len(per_batch)), key=lambda i: per_batch[i][0])
        mins = [sum(per_batch[i][c] for i in order) / nb for c in range(len(cols))]
        return avg, mins, names
