"""Device-side input pipeline: the reference's per-sample CPU transforms on a uint8 batch, as one kernel per tensor.

Reference: /root/reference/src/transform_list.py -- RandomHorizontalFlip (:161-169), RandomScaleCrop (:189-203),
ArrayToTensor (:95-113, /255), Normalize(0.5, 0.5) (:84-93); composed in src/GDN_main.py:41-66 and applied by
SequenceFolder.__getitem__ (src/datasets/datasets_list.py:81-107) to the list [gt, rgb, gt_sparse] with ONE set of
random draws per sample.  Here the draws stay on the host (same distributions, numpy RandomState), the pixels never
touch the CPU again: uint8 HWC in pinned memory -> H2D (4x fewer bytes than fp32) -> gdn_preprocess_u8.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def preprocess_u8(src, flip=None, crop=None, out=None):
    """src: uint8 CUDA tensor (N, H, W, C) or (N, H, W); flip: int32 CUDA (N,) or None; crop: float32 CUDA (N, 4) =
    (scaled_h, scaled_w, off_y, off_x) or None (the zoom is the reference's imresize on the float32 image: per-image
    min-max byte scaling + PIL's 8-bit bilinear resize, bit-exact).  Returns fp32 (N, C, H, W) in [-1, 1]."""
    if not (isinstance(src, torch.Tensor) and src.is_cuda and src.dtype == torch.uint8):
        raise RuntimeError("gdn_b200.preprocess_u8: src must be a CUDA uint8 tensor (no CPU fallback)")
    if src.dim() == 3:
        src = src.unsqueeze(-1)
    src = src.contiguous()
    N, H, W, Cc = src.shape
    if out is None:
        out = torch.empty((N, Cc, H, W), dtype=torch.float32, device=src.device)
    if flip is not None:
        flip = flip.to(device=src.device, dtype=torch.int32).contiguous()
    if crop is not None:
        crop = crop.to(device=src.device, dtype=torch.float32).contiguous()
    L = _lib.lib()
    scratch = torch.empty(2 * N, dtype=torch.int32, device=src.device) if crop is not None else None
    with torch.cuda.device(src.device):
        rc = L.gdn_preprocess_u8(C.c_void_p(src.data_ptr()), C.c_void_p(out.data_ptr()), N, H, W, Cc,
                                 C.c_void_p(flip.data_ptr() if flip is not None else None),
                                 C.c_void_p(crop.data_ptr() if crop is not None else None),
                                 C.c_void_p(scratch.data_ptr() if scratch is not None else None), _lib.stream_ptr())
    _lib.check(rc, "preprocess_u8")
    return out


class DeviceInputPipeline:
    """train=True: RandomHorizontalFlip + RandomScaleCrop + ArrayToTensor + Normalize (GDN_main.py:57-66);
    train=False: ArrayToTensor + Normalize (valid_transform, GDN_main.py:49-54).

    __call__(gt_u8, rgb_u8, sparse_u8) takes the three uint8 batches of a step ((N, H, W, C) or (N, H, W), host --
    ideally pinned -- or device) and returns (gt, rgb, gt_sparse) as fp32 (N, C, H, W) CUDA tensors in [-1, 1]: what
    the reference's DataLoader yields (trainer.py:670) after .cuda()."""

    def __init__(self, device, train=True, seed=0):
        self.device = torch.device(device)
        self.train = train
        self.rng = np.random.RandomState(seed)

    def draw(self, n, h, w):
        """the reference's random draws, per sample: flip with p = 0.5 (:165), x / y scaling ~ U(1, 1.15) (:195),
        crop offsets ~ randint (:199-200)"""
        flip = (self.rng.random_sample(n) < 0.5).astype(np.int32)
        xs, ys = self.rng.uniform(1, 1.15, n), self.rng.uniform(1, 1.15, n)
        sh, sw = (h * ys).astype(np.int64), (w * xs).astype(np.int64)
        oy = np.array([self.rng.randint(int(a) - h + 1) for a in sh])
        ox = np.array([self.rng.randint(int(a) - w + 1) for a in sw])
        crop = np.stack([sh, sw, oy, ox], 1).astype(np.float32)
        return flip, crop

    def __call__(self, gt_u8, rgb_u8, sparse_u8):
        n, h, w = rgb_u8.shape[0], rgb_u8.shape[1], rgb_u8.shape[2]
        flip = crop = None
        if self.train:
            f, c = self.draw(n, h, w)
            flip = torch.from_numpy(f).to(self.device, non_blocking=True)
            crop = torch.from_numpy(c).to(self.device, non_blocking=True)
        outs = []
        for t in (gt_u8, rgb_u8, sparse_u8):
            if t is None:
                outs.append(None)
                continue
            if not torch.is_tensor(t):
                t = torch.from_numpy(np.ascontiguousarray(t))
            outs.append(preprocess_u8(t.to(self.device, non_blocking=True), flip, crop))
        return tuple(outs)


class HostBatchPrefetcher:
    """Double-buffered host->device feed for the fused steps: while step i runs, the inputs of step i+1 are copied
    from (pinned) host memory on a dedicated copy stream, so the H2D transfer and its enqueue cost leave the step's
    critical path (the reference does ``.cuda()`` synchronously at the top of every iteration, trainer.py:686-690).

        feed = HostBatchPrefetcher(dev); feed.submit(*first_batch)
        for nxt in batches:
            x = feed.next(); feed.submit(*nxt); step.step(*x)

    Slots are recycled: a slot is overwritten only after everything that was enqueued on the consumer's stream up to
    the following ``next()`` has finished (event recorded there), i.e. after the step that read it."""

    def __init__(self, device, depth=2):
        self.dev = torch.device(device)
        self.depth = max(2, int(depth))
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.slots = [None] * self.depth
        self.ready = [None] * self.depth        # copy finished (recorded on the copy stream)
        self.consumed = [None] * self.depth     # reader finished (recorded on the consumer's stream)
        self.head = self.tail = 0               # next slot to fill / to hand out
        self._last = None

    def submit(self, *host_tensors):
        if self.head - self.tail >= self.depth:
            raise RuntimeError("HostBatchPrefetcher: %d batches already in flight" % self.depth)
        i = self.head % self.depth
        bufs = self.slots[i]
        if bufs is None or any((b is None) != (t is None) or (t is not None and (b.shape != t.shape or b.dtype != t.dtype))
                               for b, t in zip(bufs, host_tensors)) or len(bufs) != len(host_tensors):
            bufs = [None if t is None else torch.empty(t.shape, dtype=t.dtype, device=self.dev) for t in host_tensors]
            self.slots[i] = bufs
        with torch.cuda.stream(self.copy_stream):
            if self.consumed[i] is not None:
                self.copy_stream.wait_event(self.consumed[i])
            for b, t in zip(bufs, host_tensors):
                if t is not None:
                    b.copy_(t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.ready[i] = ev
        self.head += 1

    def next(self):
        if self.tail >= self.head:
            raise RuntimeError("HostBatchPrefetcher: nothing submitted")
        cur = torch.cuda.current_stream(self.dev)
        if self._last is not None:              # whatever read the previous slot has been enqueued by now
            ev = torch.cuda.Event()
            ev.record(cur)
            self.consumed[self._last] = ev
        i = self.tail % self.depth
        cur.wait_event(self.ready[i])
        self._last = i
        self.tail += 1
        return tuple(self.slots[i])
