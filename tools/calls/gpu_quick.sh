#!/bin/bash
# quick single-GPU round: GPU tests + bench (+ optional timeline).  Usage: tools/gpu_quick.sh TAG [pytest-args]
TAG=${1:-q}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q ${2:-} > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
GDN_GRAPH=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_eager.json 2> $O/${TAG}_bench_eager.err
timeout 300 python tools/timeline.py > $O/${TAG}_timeline.txt 2>&1
tail -4 $O/${TAG}_pytest.log; cat $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err; cat $O/${TAG}_bench_eager.json; head -1 $O/${TAG}_timeline.txt
