"""One RtoD training step at the bench configuration, for ncu (launch list / full capture).
   python tools/profile_step.py [steps]"""
import os, sys
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
only_last = os.environ.get("GDN_PROFILE_LAST", "0") == "1"   # with `ncu --profile-from-start off`: skip construction,
for i in range(n):                                           # autotuning and the warm-up steps
    if only_last and i == n - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    torch.cuda.nvtx.range_push("step%d" % i)
    st.step(rgb, dep, spa)
    torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
if only_last:
    torch.cuda.profiler.stop()
print("done", n, "steps; launches/step (engine fwd+bwd)", st.eng.launches_fwd + st.eng.launches_bwd)
