"""Completion-time line of one eager RtoD training step (CUDA events after every op, on the stream it ran on):
shows what overlaps what across the main / side / aux streams.   python tools/timeline.py > gpurun_out/timeline.txt"""
import os, sys
os.environ["GDN_GRAPH"] = "0"
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
for i in range(3):
    st.step(rgb, dep, spa)
torch.cuda.synchronize()
tl = []
st.eng.timeline, st.eng.tl_tag = tl, "fwd-main"
st.deng[0].timeline, st.deng[0].tl_tag = tl, "dtod-target"
st.deng[1].timeline, st.deng[1].tl_tag = tl, "dtod-pred"
e0 = torch.cuda.Event(enable_timing=True)
e1 = torch.cuda.Event(enable_timing=True)
# park the GPU (~40 ms spin) so that the host enqueues the WHOLE step before anything runs: the time line then shows
# the device's own scheduling (what a CUDA-graph replay sees), not the eager launch rate of the Python host
torch.cuda._sleep(int(float(os.environ.get("GDN_PARK_MS", "40")) * 1.9e6))
e0.record()
st.step(rgb, dep, spa)
e1.record()
torch.cuda.synchronize()
print("step %.3f ms" % e0.elapsed_time(e1))
rows = sorted(((e0.elapsed_time(ev), tag, lab) for lab, tag, ev in tl))
last = {}
for t, tag, lab in rows:
    prev = last.get(tag, 0.0)
    print("%9.3f  %-12s %-40s (+%.3f on its stream)" % (t, tag, lab, t - prev))
    last[tag] = t
