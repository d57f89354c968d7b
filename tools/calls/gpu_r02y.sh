#!/bin/bash
# round 2, call y (last GPU minutes): warp-parallel radix-select state in the metrics kernels -- whole suite (bit-exact
# delta counts, golden metrics, validation loop) and the device times of the metric kernels
TAG=${1:-r02y}; O=gpurun_out; mkdir -p $O
timeout 330 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
timeout 60 python tools/profile_metrics.py > $O/${TAG}_profile_metrics.log 2>&1; cat $O/${TAG}_profile_metrics.log
