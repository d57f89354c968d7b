#!/bin/bash
# round 2, call s: vectorised RtoD loss kernel -- tests that touch it, timing against the scalar kernel's 0.0497 ms
TAG=${1:-r02s}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py tests/test_gpu_guidance_grad.py -m gpu -x -q -k "loss or step or guidance or golden" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
timeout 200 python tools/profile_metrics.py > $O/${TAG}_profile_metrics.log 2>&1; cat $O/${TAG}_profile_metrics.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE-OK')" 2>&1 | tail -2
