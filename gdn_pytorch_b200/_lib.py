"""ctypes binding of libgdn_b200.so (the C ABI declared in include/gdn_b200.h).

The library is built in-tree by `make` / `__graft_entry__.build()`.  There is no fallback: if the shared
object is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgdn_b200.so")


class Act(C.Structure):
    """gdn_act: NHWC bf16 activation with a physical border"""
    _fields_ = [("ptr", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("pad", C.c_int32)]


class ConvDesc(C.Structure):
    _fields_ = [
        ("src0", Act), ("src1", Act), ("weights", C.c_void_p),
        ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32),
        ("off_y", C.c_int32), ("off_x", C.c_int32), ("out_h", C.c_int32), ("out_w", C.c_int32),
        ("cout", C.c_int32), ("cout_pad", C.c_int32), ("algo", C.c_int32),
        ("bias", C.c_void_p), ("relu", C.c_int32), ("tanh_out", C.c_int32),
        ("resid", C.c_void_p), ("out_f32", C.c_void_p), ("out_bf16", Act), ("out_reflect", C.c_int32),
        ("dst_h", C.c_int32), ("dst_w", C.c_int32), ("dst_sy", C.c_int32), ("dst_sx", C.c_int32),
        ("dst_oy", C.c_int32), ("dst_ox", C.c_int32),
        ("stat_sum", C.c_void_p), ("stat_sqsum", C.c_void_p), ("out16_is_half", C.c_int32),
        ("bwd_raw", C.c_void_p), ("bwd_coef", C.c_void_p), ("bwd_relu", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("fin_counter", C.c_void_p), ("fin_gamma", C.c_void_p), ("fin_beta", C.c_void_p), ("fin_running_mean", C.c_void_p),
        ("fin_running_var", C.c_void_p), ("fin_scale", C.c_void_p), ("fin_shift", C.c_void_p), ("fin_mean", C.c_void_p),
        ("fin_rstd", C.c_void_p), ("fin_coef4", C.c_void_p), ("fin_count", C.c_double), ("fin_eps", C.c_float),
        ("fin_momentum", C.c_float),
    ]


class WgradDesc(C.Structure):
    _fields_ = [
        ("x0", Act), ("x1", Act), ("dy", Act), ("dw", C.c_void_p),
        ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("off_y", C.c_int32), ("off_x", C.c_int32),
        ("out_h", C.c_int32), ("out_w", C.c_int32), ("cout_pad", C.c_int32),
        ("slabs", C.c_void_p), ("max_slabs", C.c_int32), ("splits_used", C.POINTER(C.c_int32)),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "gdn_b200: %s not found -- build it with `make` (or __graft_entry__.build()); "
                "there is no CPU / PyTorch fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.gdn_last_error.restype = C.c_char_p
        L.gdn_version.restype = C.c_int
        L.gdn_sm_count.restype = C.c_int
        L.gdn_resize_u8_workspace.restype = C.c_size_t
        L.gdn_conv2d_workspace_bytes.restype = C.c_size_t
        L.gdn_depth_metrics_workspace_bytes.restype = C.c_size_t
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError("gdn_b200 %s failed (%d): %s" % (what, rc, lib().gdn_last_error().decode()))


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
