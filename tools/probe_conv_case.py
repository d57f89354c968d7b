"""One shape of tools/sweep_conv.py with one algo word, launched a few times: the target of an `ncu --set full` capture
(development aid: where does a short-reduction convolution spend its time?).
    python tools/probe_conv_case.py <substring of the shape title> <algo hex> [launches]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.argv, args = sys.argv[:1], sys.argv[1:]
import sweep_conv as S
import torch

title, shape = [(t, a) for t, a in S.SHAPES if args[0] in t][0]
d, flops, keep = S.make(*shape)
d.algo = int(args[1], 16)
n = int(args[2]) if len(args) > 2 else 4
for i in range(n):
    rc = S.L.gdn_conv2d(C.byref(d), S._lib.stream_ptr())
    assert rc == 0, S.L.gdn_last_error()
torch.cuda.synchronize()
ms = S.bench(d)
print("%s  %s: %.4f ms  %.0f TFLOP/s" % (title, S.name(d.algo), ms, flops / ms / 1e9))
