"""The dominant kernel alone (64->64 k9 implicit-GEMM conv at the bench shape), for `ncu --set full`:
its dram__bytes_read/write per launch is bench.py's roofline.traffic (profiles/dominant_conv_traffic.json).
The autotune candidates and warm-up launches run BEFORE cudaProfilerStart: capture with
    ncu --profile-from-start off --set full --clock-control none -k regex:conv_igemm -c 2 python tools/profile_conv.py [B]
so that only the SHIPPED variant is in the report (round 1's `-s 3 -c 2` landed on autotune candidates)."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ms, fl, algo = bench.time_dominant_conv(dev, B, profile_launches=3)
print("conv 64->64 k9 B=%d: %.4f ms  %.1f TFLOP/s  algo 0x%x" % (B, ms, fl / ms / 1e9, algo))
