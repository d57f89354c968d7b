#!/bin/bash
# round 2, call w: ncu --set full (source counters) of the 128->64 1x1 convolution AFTER the epilogue work (eight epilogue
# warps, wide pixel tiles, saturating pack) -- the counterpart of profiles/r02l_ncu_conv1x1.txt
TAG=${1:-r02w}; O=gpurun_out; mkdir -p $O
NCU="ncu --clock-control none --set full --import-source on -k regex:conv_igemm -s 4 -c 1 -f"
timeout 300 $NCU -o $O/${TAG}_conv1x1_concat python tools/probe_conv_case.py "128->64 k1" 0x11010001 > $O/${TAG}_ncu_conv1x1.log 2>&1; tail -2 $O/${TAG}_ncu_conv1x1.log
python tools/ncu_summary.py $O/${TAG}_conv1x1_concat.ncu-rep > $O/${TAG}_ncu_conv1x1.summary.txt 2>&1; cat $O/${TAG}_ncu_conv1x1.summary.txt
