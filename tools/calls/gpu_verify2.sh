#!/bin/bash
# last call of the round (GPU budget ~100 s): the tests that exercise the two kernels changed since the last full
# green run (fold_rows2, wide-tile pack), then a short bench
TAG=${1:-r01w}; O=gpurun_out; mkdir -p $O
timeout 70 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_network.py tests/test_gpu_new_kernels.py tests/test_gpu_guidance_grad.py -x -q -m gpu > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
timeout 45 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
cat $O/${TAG}_bench.json
