#!/bin/bash
# short verification call: new tests (all failures shown), whole GPU suite, default bench, smoke
TAG=${1:-r01s}; O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_new_kernels.py tests/test_gpu_data.py -q -m gpu > $O/${TAG}_pytest_new.log 2>&1; echo "rc=$?" >> $O/${TAG}_pytest_new.log
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1
tail -25 $O/${TAG}_pytest_new.log; tail -4 $O/${TAG}_pytest.log; cat $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err; tail -2 $O/${TAG}_smoke.log
