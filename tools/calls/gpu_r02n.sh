#!/bin/bash
# round 2, call n: deterministic mode (bit-identical runs + its cost), wide pixel tiles of the 1x1 launches, saturating pack
TAG=${1:-r02n}; O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
GDN_DETERMINISTIC=1 timeout 600 python tools/check_deterministic.py 10 4 > $O/${TAG}_deterministic.log 2>&1; tail -6 $O/${TAG}_deterministic.log
timeout 300 python tools/check_deterministic.py 4 4 > $O/${TAG}_nondeterministic.log 2>&1; tail -5 $O/${TAG}_nondeterministic.log
B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 20 --warmup 5"
timeout 300 $B > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 300 $B > $O/${TAG}_bench2.json 2>> $O/${TAG}_bench.err
GDN_DETERMINISTIC=1 timeout 300 $B > $O/${TAG}_bench_det.json 2>> $O/${TAG}_bench.err
timeout 300 $B --workload infer > $O/${TAG}_bench_infer.json 2>> $O/${TAG}_bench.err
for f in bench bench2 bench_det bench_infer; do python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_$f.json")); print("%-16s %7.1f img/s  %6.2f ms  e2e %7.1f" % ("$f", d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$f: no result", e)
PY
done
timeout 300 python tools/profile_ops.py > $O/${TAG}_profile_ops.log 2>&1; grep "1x1\|upconv4\|head\|downconv0\|FORWARD\|BACKWARD\|DTOD\|sum of\|by kind" $O/${TAG}_profile_ops.log
grep -v "OMP_NUM\|\*\*\*\*\|^$" $O/${TAG}_bench.err | tail -5 | cut -c1-300
