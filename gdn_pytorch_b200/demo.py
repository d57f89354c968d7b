"""Demo / single-image inference path: /root/reference/src/depth_extract.py:60-147 on the device.

The reference loads an image, resizes it on the CPU to 128x416 with ``scipy.misc.imresize`` (:86, class Resize
:23-58), normalises it (:90-91), runs ``AutoEncoder`` at batch 1 (:122), resizes the depth map back to the original
size with ``imresize`` again (:138) and writes it as an 8-bit image (:145).  Here the uint8 image goes to the GPU
once; bytescale, PIL-exact bilinear resizing, normalisation, the network (ONE captured CUDA graph at batch 1 -- the
~60 kernel launches of the inference plan become one cudaGraphLaunch, which is what bounds B=1 latency) and the
resize back all run there, and the 8-bit depth image comes back.

``imresize`` / ``bytescale`` are drop-ins for the scipy.misc functions the reference calls (removed from SciPy 1.3)
on CUDA tensors, bit-exact w.r.t. PIL's 8-bit BILINEAR resize.  No CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .data import preprocess_u8
from .engine import Engine
from .module_runtime import _params, _check_norm


def _need_cuda(t, name, dtypes):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype in dtypes):
        raise RuntimeError("gdn_b200.%s: expected a CUDA tensor of dtype %s (no CPU fallback)" % (name, dtypes))


def bytescale(t, out=None, f64=False):
    """scipy.misc.bytescale with its defaults: min-max stretch to uint8.  t: CUDA float32 or uint8 tensor (uint8 is
    read as the float image it stores, as depth_extract.py's load_as_float does).  f64: evaluate in double precision
    like NumPy does for a float64 array (the demo's output path, depth_extract.py:135-138)."""
    _need_cuda(t, "bytescale", (torch.float32, torch.uint8))
    t = t.contiguous()
    if out is None:
        out = torch.empty(t.shape, dtype=torch.uint8, device=t.device)
    scratch = torch.empty(2, dtype=torch.int32, device=t.device)
    with torch.cuda.device(t.device):
        rc = _lib.lib().gdn_bytescale(C.c_void_p(t.data_ptr()), int(t.dtype == torch.uint8), C.c_int64(t.numel()),
                                      int(bool(f64)), C.c_void_p(out.data_ptr()), C.c_void_p(scratch.data_ptr()), _lib.stream_ptr())
    _lib.check(rc, "bytescale")
    return out


def resize_u8(t, size, out=None):
    """PIL.Image.resize(size[::-1], BILINEAR) of uint8 CUDA images: t (H, W), (H, W, C) or (N, H, W, C); size = (h, w)"""
    _need_cuda(t, "resize_u8", (torch.uint8,))
    shp = tuple(t.shape)
    if t.dim() == 2:
        t4 = t.reshape(1, shp[0], shp[1], 1)
    elif t.dim() == 3:
        t4 = t.reshape(1, *shp)
    elif t.dim() == 4:
        t4 = t
    else:
        raise ValueError("resize_u8: expected (H,W), (H,W,C) or (N,H,W,C)")
    t4 = t4.contiguous()
    n, h, w, c = t4.shape
    oh, ow = int(size[0]), int(size[1])
    L = _lib.lib()
    ws_bytes = L.gdn_resize_u8_workspace(n, h, w, c, oh, ow)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=t.device)
    if out is None:
        out = torch.empty((n, oh, ow, c), dtype=torch.uint8, device=t.device)
    with torch.cuda.device(t.device):
        rc = L.gdn_resize_u8(C.c_void_p(t4.data_ptr()), C.c_void_p(out.data_ptr()), n, h, w, c, oh, ow,
                             C.c_void_p(ws.data_ptr()), C.c_size_t(ws_bytes), _lib.stream_ptr())
    _lib.check(rc, "resize_u8")
    if t.dim() == 2:
        return out.reshape(oh, ow)
    if t.dim() == 3:
        return out.reshape(oh, ow, c)
    return out


def imresize(t, size, f64=False):
    """scipy.misc.imresize(arr, size, 'bilinear') (depth_extract.py:41-58) for CUDA tensors: (H,W) or (H,W,3),
    float32 or uint8-stored float image; size = (h, w).  Returns uint8."""
    return resize_u8(bytescale(t, f64=f64), size)


class DepthExtractor:
    """``extractor(img)`` = the body of the demo loop (depth_extract.py:84-91 and :114-141) for one image.

    img: uint8 (H, W, 3) image -- numpy array, host tensor (ideally pinned) or CUDA tensor.
    Returns the 8-bit depth image (org_H, org_W) as a CUDA uint8 tensor (what ``imsave`` writes, :145);
    ``last_depth`` keeps the network output (1, 1, height, width) fp32 of the last call."""

    def __init__(self, model, height=128, width=416, use_graph=True):
        m = model.module if hasattr(model, "module") else model      # nn.DataParallel wrapper (depth_extract.py:66)
        if not hasattr(m, "gdn_graph"):
            raise TypeError("DepthExtractor needs a gdn_pytorch_b200.AE_model_unet network")
        if m.training:
            raise RuntimeError("DepthExtractor: call model.eval() first (depth_extract.py:68)")
        _check_norm(m)
        self.model = m
        self.dev = next(m.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("gdn_b200: CUDA-only (sm_100a) implementation; move the model to the GPU")
        self.h, self.w = int(height), int(width)
        with torch.cuda.device(self.dev):
            self.eng = Engine(m.gdn_graph(), _params(m), 1, self.h, self.w, train=False, want=(), device=self.dev)
            self.static_in = torch.zeros((1, 3, self.h, self.w), dtype=torch.float32, device=self.dev)
            self.small_u8 = torch.empty((1, self.h, self.w, 3), dtype=torch.uint8, device=self.dev)
            self.depth_u8 = torch.empty((self.h, self.w), dtype=torch.uint8, device=self.dev)
        self.use_graph = use_graph
        self._graph = None
        self._warm = 0
        self.last_depth = None

    def _network(self):
        eng = self.eng
        eng.refresh_if_stale()                 # weight (re)packing stays outside the captured graph
        if not self.use_graph:
            eng.forward(self.static_in)
            return
        if self._graph is None:
            if self._warm < 1:                 # one eager pass first (lazy module loading, allocator warm-up)
                eng.forward(self.static_in)
                self._warm += 1
                return
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                eng.forward(self.static_in)
            self._graph = g
        self._graph.replay()

    @torch.no_grad()
    def __call__(self, img):
        if isinstance(img, np.ndarray):
            img = torch.from_numpy(np.ascontiguousarray(img))
        if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3:
            raise ValueError("DepthExtractor: expected a uint8 (H, W, 3) image")
        with torch.cuda.device(self.dev):
            img = img.to(self.dev, non_blocking=True)
            org_h, org_w = int(img.shape[0]), int(img.shape[1])
            resize_u8(bytescale(img), (self.h, self.w), out=self.small_u8)            # Resize()(img, (128, 416)) :86
            preprocess_u8(self.small_u8, out=self.static_in)                          # /255, (x - 0.5) / 0.5   :88-91
            self._network()                                                           # ae(tens, istrain=False) :122
            depth = self.eng.depth()
            self.last_depth = depth
            bytescale(depth.reshape(self.h, self.w), out=self.depth_u8, f64=True)     # imresize -> toimage   :138
            return resize_u8(self.depth_u8, (org_h, org_w))
