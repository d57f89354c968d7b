#!/bin/bash
# round 2, call g: staged Eigen-metrics kernels (tests + timing), inference workloads, full suite
TAG=${1:-r02g}; O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -6 $O/${TAG}_pytest.log
timeout 200 python tools/profile_metrics.py > $O/${TAG}_profile_metrics.log 2>&1; cat $O/${TAG}_profile_metrics.log
timeout 300 python bench.py --workload infer --steps 20 --warmup 3 > $O/${TAG}_bench_infer.json 2> $O/${TAG}_bench.err
timeout 300 python bench.py --workload infer_fullres --steps 10 --warmup 3 > $O/${TAG}_bench_infer_fullres.json 2>> $O/${TAG}_bench.err
timeout 300 python bench.py --workload train_dtod --steps 10 --warmup 3 > $O/${TAG}_bench_train_dtod.json 2>> $O/${TAG}_bench.err
for f in infer infer_fullres train_dtod; do python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_bench_$f.json")); print("%-14s %7.1f img/s  %6.2f ms  e2e %7.1f" % ("$f", d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$f: no result", e)
PY
done
tail -5 $O/${TAG}_bench.err
