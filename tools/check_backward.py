"""GPU check of engine forward+backward (train mode) on shallow graphs against torch autograd through the fp32
graph interpreter.   python tools/check_backward.py [-v]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from gdn_pytorch_b200.engine import Engine
from oracle import synth
from oracle.graph_interp import run_graph
from tests import minigraphs

dev = "cuda"


def rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def run(gname, B=2, H=32, W=64, verbose=False):
    g = getattr(minigraphs, gname)()
    sd = minigraphs.synth_params(g, 0, dev)
    for k, v in sd.items():
        if not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
    x = (synth.synth_rgb(B, H, W, 1) if g.cin == 3 else synth.synth_depth(B, H, W, 1)).to(dev)
    R = (torch.rand((B, 1, H, W), generator=torch.Generator().manual_seed(5)) - 0.5).to(dev)
    # engine forward first: its ReLU masks are imposed on the reference so that the comparison is not dominated by
    # sign flips of near-zero pre-activations
    names = [u.out for u in g.units]
    eng = Engine(g, sd, B, H, W, train=True, backward=True, want=names)
    eng.forward(x)
    masks = {u.out: (eng.value_nchw(u.out) > 0) for u in g.units if u.relu and not u.resid}
    # reference (fp64 for clean gradients)
    sd64 = {k: v.detach().double().requires_grad_(v.requires_grad) for k, v in sd.items()}
    T = run_graph(g, sd64, x.double(), train=True, relu_masks=masks)
    for n_, t_ in T.items():
        if t_.requires_grad:
            t_.retain_grad()
    loss = (T["out"] * R.double()).sum()
    pn = [k for k, v in sd64.items() if v.requires_grad]
    loss.backward()
    gref = {k: sd64[k].grad for k in pn}
    out = eng.depth()
    dpre = R * (1 - out * out)
    eng.flat_grad.zero_()
    eng.backward(dpre.view(B, H, W))
    torch.cuda.synchronize()
    worst_f = max(rel(eng.value_nchw(n), T[n].float()) for n in names)
    print("%s: forward worst tensor err %.2e (out %.2e)" % (gname, worst_f, rel(out, T["out"].float())))
    def l2(a, b):
        return ((a - b).norm() / (b.norm() + 1e-30)).item()
    for u in reversed(g.units):
        if u.out in eng.dact and T[u.out].grad is not None and u.out != "out":
            a = eng.dact[u.out].permute(0, 3, 1, 2)
            b = T[u.out].grad.float()
            print("   dact %-18s max %.2e  l2 %.2e" % (u.out, rel(a, b), l2(a, b)))
    bad = 0
    for k in pn:
        e = rel(eng.grad[k], gref[k].float())
        if gref[k].abs().max().item() < 1e-9:
            e = 0.0   # analytically-zero gradient (per-channel shift in front of a batch-stat BN)
        e2 = l2(eng.grad[k], gref[k].float()) if gref[k].abs().max().item() >= 1e-9 else 0.0
        flag = "" if e2 < 5e-2 else "  <<<<<<"
        k = k + " l2=%.1e" % e2
        bad += bool(flag)
        if verbose or flag:
            print("   grad %-40s err %.2e%s" % (k, e, flag))
    print("%s: %d/%d parameter gradients within 5e-2; launches fwd %d bwd %d" % (gname, len(pn) - bad, len(pn),
                                                                                 eng.launches_fwd, eng.launches_bwd), flush=True)
    return bad == 0


if __name__ == "__main__":
    v = "-v" in sys.argv
    ok = True
    for gname in ("mini_rtod", "mini_dtod", "mini_deep512"):
        ok &= run(gname, verbose=v)
    print("ALL OK" if ok else "SOME FAILED")
