#!/bin/bash
# round 2, call b: conv variant sweep on the slow shapes + train-mode parity probe at the headline shape
TAG=${1:-r02b}; O=gpurun_out; mkdir -p $O
timeout 600 python tools/sweep_conv.py > $O/${TAG}_sweep_conv.log 2>&1
timeout 900 python -m tests.parity_probe fwd grad --b 4 > $O/${TAG}_parity_probe.log 2>&1
timeout 900 python -m tests.parity_probe traj --b 4 --steps 50 --kinds init >> $O/${TAG}_parity_probe.log 2>&1
tail -60 $O/${TAG}_parity_probe.log
