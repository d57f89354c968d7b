#!/bin/bash
# round 2, call f: graph-timed autotune / sweep / per-op table (true device times), new pack + upsample kernels, tests, bench
TAG=${1:-r02f}; O=gpurun_out; mkdir -p $O
B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 10 --warmup 3"
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -6 $O/${TAG}_pytest.log
timeout 600 python tools/sweep_conv.py > $O/${TAG}_sweep_conv.log 2>&1
timeout 300 $B > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench.err
GDN_SPLITK=0 timeout 300 $B > $O/${TAG}_bench_nosplit.json 2>> $O/${TAG}_bench.err
timeout 300 $B > $O/${TAG}_bench_default2.json 2>> $O/${TAG}_bench.err
timeout 400 python tools/profile_ops.py > $O/${TAG}_profile_ops.log 2>&1
for f in default nosplit default2; do python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_bench_$f.json")); print("%-10s %7.1f img/s  %6.2f ms  e2e %7.1f" % ("$f", d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$f: no result", e)
PY
done
grep -E "by kind|====|sum of ops" $O/${TAG}_profile_ops.log; grep -A5 "8x26\|16x52 train\|s2" $O/${TAG}_sweep_conv.log | head -80; tail -5 $O/${TAG}_bench.err
