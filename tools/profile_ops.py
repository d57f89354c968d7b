"""Per-op device times of one RtoD training step at the bench configuration (CUDA events, each op repeated)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gdn_pytorch_b200.trainer import RtoDTrainStep

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
B = int(os.environ.get("GDN_BATCH", "20"))
rgb, dep, spa = [t.to(dev) for t in bench.synth_batch(B, 0)]
rtod, dtod = bench.build_models(dev)
st = RtoDTrainStep(rtod, dtod)
for i in range(2):
    st.step(rgb, dep, spa)
torch.cuda.synchronize()
eng = st.eng
tot = 0
for title, ops in (("FORWARD", eng.fwd), ("BACKWARD", eng.bwd), ("DTOD-ENC", st.deng[0].fwd)):
    res = eng.profile(ops)
    t = sum(ms for _, ms in res)
    tot += t * (2 if title == "DTOD-ENC" else 1)
    print("==== %s  %.3f ms" % (title, t))
    for n, ms in res:
        if ms >= 0.02:
            print("   %-44s %8.3f ms" % (n, ms))
    agg = collections.defaultdict(float)
    for n, ms in res:
        agg[n.split(" ")[0]] += ms
    print("   by kind:", ", ".join("%s %.2f" % kv for kv in sorted(agg.items(), key=lambda kv: -kv[1])))
print("sum of ops: %.2f ms" % tot)
