"""GPU check of the engine against the fp32 graph interpreter / oracle, tensor by tensor.
   python tools/check_engine.py [eval|train|bwd|time] ..."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from gdn_pytorch_b200 import graph as G
from gdn_pytorch_b200.engine import Engine
from oracle import synth
from oracle.graph_interp import run_graph
from tests.util import shapes_of, build_module

dev = "cuda"


def rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-20)).item()


def check_forward(name, cin, train, B=2, H=32, W=64, verbose=True, bf16_ref=True):
    sd = {k: v.to(dev) for k, v in synth.synth_state_dict(shapes_of(name), seed=0).items()}
    for k in sd:
        if sd[k].dtype == torch.float32 and not k.endswith(("running_mean", "running_var")):
            sd[k].requires_grad_(False)
    x = (synth.synth_rgb(B, H, W, 0) if cin == 3 else synth.synth_depth(B, H, W, 0)).to(dev)
    g = G.GRAPHS[name]() if name == "AutoEncoder" else G.GRAPHS[name](cin)
    names = [u.out for u in g.units]
    sd_ref = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        T32 = run_graph(g, sd_ref, x, train=train, bf16=False)
        Tbf = run_graph(g, sd_ref, x, train=train, bf16=True)
    eng = Engine(g, sd, B, H, W, train=train, backward=False, want=names)
    eng.forward(x)
    torch.cuda.synchronize()
    worst = 0
    for u in g.units:
        got = eng.value_nchw(u.out)
        e32, ebf = rel(got, T32[u.out]), rel(got, Tbf[u.out])
        flag = "" if ebf < 2e-2 else "   <<<<<<"
        worst = max(worst, ebf)
        if verbose or flag:
            print("  %-22s %-18s vs fp32 %.2e  vs bf16-emul %.2e%s" % (u.conv, u.out, e32, ebf, flag))
    out = g.units[-1].out
    print("%s %s B%d %dx%d: depth err vs fp32 %.3e, vs bf16-emul %.3e ; worst tensor vs bf16-emul %.2e ; launches %d" % (
        name, "train" if train else "eval", B, H, W, rel(eng.value_nchw(out), T32[out]), rel(eng.value_nchw(out), Tbf[out]),
        worst, eng.launches_fwd), flush=True)
    return eng


def main():
    what = sys.argv[1:] or ["eval"]
    nets = [("AutoEncoder_2", 3), ("AutoEncoder_DtoD", 1), ("AutoEncoder", 3)]
    if "eval" in what:
        for name, cin in nets:
            check_forward(name, cin, False, verbose="-v" in what)
    if "train" in what:
        for name, cin in nets[:2]:
            check_forward(name, cin, True, verbose="-v" in what)
    if "full" in what:
        for name, cin in nets:
            check_forward(name, cin, False, B=2, H=128, W=416, verbose=False)


if __name__ == "__main__":
    main()
