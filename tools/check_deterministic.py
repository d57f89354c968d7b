"""GDN_DETERMINISTIC=1: are two runs of the same training steps bit-identical?  (single GPU)
Runs K fused RtoD steps twice eagerly and twice as a CUDA graph in this process (fresh models and engines every time) and
compares them all gradients and parameters bit for bit (the training state) and the reported loss values to 1e-9 relative: the
loss SUMS are fp64 atomics over per-CTA fp64 partials -- they only feed the value that is reported, never a gradient -- so
their last bit may depend on the order.  Prints DET-OK / DET-FAIL.
    GDN_DETERMINISTIC=1 python tools/check_deterministic.py [steps] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gdn_pytorch_b200 import _lib
from gdn_pytorch_b200.trainer import RtoDTrainStep

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
LR = 2e-5
det = bool(_lib.lib().gdn_deterministic())
print("library deterministic mode: %s" % det)


def run(graph):
    os.environ["GDN_GRAPH"] = "1" if graph else "0"
    rtod, dtod = bench.build_models(dev)
    st = RtoDTrainStep(rtod, dtod, lr=LR)
    batches = [[t.to(dev) for t in bench.synth_batch(B, i)] for i in range(2)]
    losses, grads = [], []
    for i in range(STEPS):
        losses.append(st.step(*batches[i % 2])["loss"].clone())
        grads.append(st.eng.flat_grad.clone())
    torch.cuda.synchronize()
    return st.flat_params.clone(), torch.stack([l.reshape(()) for l in losses]), grads


def compare(name, a, b):
    p, l, g = a
    same_l = bool(((l.double() - b[1].double()).abs() <= 1e-9 * b[1].double().abs()).all())
    first_bad = next((i for i, (x, y) in enumerate(zip(g, b[2])) if not torch.equal(x, y)), None)
    same_p = torch.equal(p, b[0])
    print("%-28s losses equal to 1e-9 %s (bitwise %s) | gradients identical %s | parameters identical %s (max |dp| %.3g)"
          % (name, same_l, torch.equal(l, b[1]), "yes" if first_bad is None else "first differ at step %d" % first_bad, same_p,
             (p - b[0]).abs().max().item()))
    return same_l and same_p and first_bad is None


e1, e2 = run(False), run(False)
g1, g2 = run(True), run(True)
ok = compare("eager vs eager", e2, e1)
ok = compare("CUDA graph vs CUDA graph", g2, g1) and ok
ok = compare("CUDA graph vs eager", g1, e1) and ok
print("losses:", ["%.7f" % v for v in e1[1].tolist()])
print(("DET-OK" if ok else "DET-FAIL") + " steps=%d batch=%d deterministic=%s" % (STEPS, B, det))
sys.exit(0 if (ok or not det) else 1)
