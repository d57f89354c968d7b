#!/bin/bash
# round 2, call q: whole suite after the shared finalisation arithmetic / reverted unpack clear; default bench
TAG=${1:-r02q}; O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 20 --warmup 5"
timeout 300 $B > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 300 $B --workload infer > $O/${TAG}_bench_infer.json 2>> $O/${TAG}_bench.err
for f in bench bench_infer; do python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_$f.json")); print("%-16s %7.1f img/s  %6.2f ms  e2e %7.1f  launches/step %d" % ("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"] / d["steps"]))
except Exception as e:
    print("$f: no result", e)
PY
done
grep -v "OMP_NUM\|\*\*\*\*\|^$" $O/${TAG}_bench.err | tail -5 | cut -c1-300
