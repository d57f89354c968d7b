"""The kernels of the 8f rows added last (demo path, guidance gradient), eagerly, for an `ncu -k regex:...` capture:
   python tools/profile_extras.py"""
import os, sys
os.environ["GDN_GRAPH"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import contextlib, io
import numpy as np
import torch
import bench
from gdn_pytorch_b200 import AE_model_unet as M
from gdn_pytorch_b200.demo import DepthExtractor
from gdn_pytorch_b200.trainer import RtoDTrainStep

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
with contextlib.redirect_stdout(io.StringIO()):
    torch.manual_seed(0)
    ae = M.AutoEncoder(height=128, width=416).to(dev).eval()
ex = DepthExtractor(ae, use_graph=False)
frame = torch.from_numpy(np.random.RandomState(0).randint(0, 256, (375, 1242, 3)).astype(np.uint8)).to(dev)
for _ in range(2):
    out = ex(frame)
torch.cuda.synchronize()
print("demo ok", tuple(out.shape))
B = int(os.environ.get("GDN_BATCH", "20"))
rgb, dep, spa = [t.to(dev) for t in bench.synth_batch(B, 0)]
rtod, dtod = bench.build_models(dev)
st = RtoDTrainStep(rtod, dtod, guidance_grad=True)
for _ in range(2):
    terms = st.step(rgb, dep, spa)
torch.cuda.synchronize()
print("guided ok", float(terms["loss"]))
