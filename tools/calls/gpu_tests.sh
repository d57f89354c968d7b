#!/bin/bash
# kernel-level GPU tests first (no -x: show every failing case), then the network tests, then a short bench
O=gpurun_out; TAG=${1:-t}; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_kernels.py -q -m gpu 2>&1 | tail -40 > $O/${TAG}_kernels.log
timeout 900 python -m pytest tests/test_gpu_network.py -q -m gpu 2>&1 | tail -40 > $O/${TAG}_network.log
timeout 300 python tools/profile_ops.py > $O/${TAG}_profile_ops.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
cat $O/${TAG}_kernels.log | tail -25; tail -25 $O/${TAG}_network.log; cat $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err; grep -E "^====|by kind|sum of" $O/${TAG}_profile_ops.log
