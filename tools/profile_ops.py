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
    res4 = eng.profile(ops, with_flops=True)
    res = [(n, ms) for n, ms, _, _ in res4]
    t = sum(ms for _, ms in res)
    tot += t * (2 if title == "DTOD-ENC" else 1)
    print("==== %s  %.3f ms" % (title, t))
    for n, ms, fl, tiles in res4:
        if ms >= 0.02:
            extra = ""
            if fl:      # tensor-core launch: achieved rate, and how many 128-pixel M tiles feed the 148 SMs
                extra = "  %7.1f GFLOP %7.0f TFLOP/s" % (fl / 1e9, fl / (ms * 1e-3) / 1e12) + ("  %5d tiles" % tiles if tiles else "")
            print("   %-44s %8.3f ms%s" % (n, ms, extra))
    agg = collections.defaultdict(float)
    for n, ms in res:
        agg[n.split(" ")[0]] += ms
    print("   by kind:", ", ".join("%s %.2f" % kv for kv in sorted(agg.items(), key=lambda kv: -kv[1])))
    # tensor-core launches grouped by achieved rate: where the FLOPs run slowly
    tc = [(fl, ms) for _, ms, fl, _ in res4 if fl]
    if tc:
        tf, tm = sum(f for f, _ in tc), sum(m for _, m in tc)
        print("   tensor-core launches: %.2f TFLOP in %.2f ms = %.0f TFLOP/s" % (tf / 1e12, tm, tf / (tm * 1e-3) / 1e12))
        for lo, hi in ((0, 400), (400, 800), (800, 1200), (1200, 9999)):
            sel = [(f, m) for f, m in tc if lo <= f / (m * 1e-3) / 1e12 < hi]
            if sel:
                print("      %4d-%4d TFLOP/s: %3d launches %6.2f ms %6.2f TFLOP" % (lo, hi, len(sel), sum(m for _, m in sel),
                                                                                   sum(f for f, _ in sel) / 1e12))
print("sum of ops: %.2f ms" % tot)
