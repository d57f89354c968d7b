#!/bin/bash
# minimal verification: the exact commands the driver runs at round end (GPU suite with -x, smoke)
TAG=${1:-v}; O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -6 $O/${TAG}_pytest.log
