"""How far apart do two runs of the same K training steps end up?  (single GPU; development aid)
Pairs compared: eager vs eager, eager vs CUDA-graph, each with and without the side streams.  Differences at the
fp32-atomic noise level are expected; anything larger is a missing stream dependency."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gdn_pytorch_b200.trainer import RtoDTrainStep

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
LR, STEPS, B = 1e-6, 4, int(os.environ.get("GDN_BATCH", "4"))


def run(graph, side):
    os.environ["GDN_GRAPH"] = "1" if graph else "0"
    os.environ["GDN_SIDE"] = "1" if side else "0"
    rtod, dtod = bench.build_models(dev)
    st = RtoDTrainStep(rtod, dtod, lr=LR)
    rgb, dep, spa = [t.to(dev) for t in bench.synth_batch(B, 0)]
    losses, grads = [], []
    for i in range(STEPS):
        losses.append(st.step(rgb, dep, spa)["loss"].clone())
        grads.append(st.eng.flat_grad.clone())
    torch.cuda.synchronize()
    return st.flat_params.clone(), [float(v) for v in losses], grads


ref_p, ref_l, ref_g = run(False, False)
print("reference: eager, no side streams; losses", ["%.7f" % v for v in ref_l])
for name, graph, side in (("eager/noside (repeat)", False, False), ("eager/side", False, True), ("graph/noside", True, False),
                          ("graph/side", True, True)):
    p, l, g = run(graph, side)
    d = (p - ref_p).abs()
    gd = [((a - b).abs().max().item() / (b.abs().max().item() + 1e-30)) for a, b in zip(g, ref_g)]
    print("%-22s |dp| max %.3g mean %.3g (lr*steps = %.3g)  grad rel-max-diff per step %s  losses %s" %
          (name, d.max().item(), d.mean().item(), LR * STEPS, ["%.2g" % v for v in gd], ["%.7f" % v for v in l]))
