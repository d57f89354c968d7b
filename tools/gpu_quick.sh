#!/bin/bash
# quick single-GPU round: GPU tests + bench (side streams on / off).  Usage: tools/gpu_quick.sh TAG
TAG=${1:-q}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
GDN_SIDE=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_noside.json 2> $O/${TAG}_bench_noside.err
tail -4 $O/${TAG}_pytest.log; cat $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err; cat $O/${TAG}_bench_noside.json
