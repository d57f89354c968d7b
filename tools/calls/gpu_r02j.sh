#!/bin/bash
# round 2, call j: whole GPU suite on the current head + default bench (with the library GPU baseline leg)
TAG=${1:-r02j}; O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -6 $O/${TAG}_pytest.log
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-1500 $O/${TAG}_bench.json
tail -3 $O/${TAG}_bench.err
