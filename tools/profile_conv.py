"""The dominant kernel alone (64->64 k9 implicit-GEMM conv at the bench shape), for `ncu --set full`:
its dram__bytes_read/write per launch is bench.py's roofline.traffic.   python tools/profile_conv.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ms, fl, algo = bench.time_dominant_conv(dev, B)
print("conv 64->64 k9 B=%d: %.4f ms  %.1f TFLOP/s  algo 0x%x" % (B, ms, fl / ms / 1e9, algo))
