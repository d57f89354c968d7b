#!/bin/bash
# round 2, call o: whole suite (deterministic mode per launch mode, wide 1x1 tiles, eight epilogue warps everywhere),
# A/B of the side-stream forward re-pack with the table kernel, device timeline of one step
TAG=${1:-r02o}; O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
GDN_DETERMINISTIC=1 timeout 600 python tools/check_deterministic.py 10 4 > $O/${TAG}_deterministic.log 2>&1; tail -6 $O/${TAG}_deterministic.log
B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 20 --warmup 5"
timeout 300 $B > $O/${TAG}_bench_syncpack.json 2> $O/${TAG}_bench.err
GDN_ASYNC_FWD_PACK=1 timeout 300 $B > $O/${TAG}_bench_asyncpack.json 2>> $O/${TAG}_bench.err
timeout 300 $B > $O/${TAG}_bench_syncpack2.json 2>> $O/${TAG}_bench.err
GDN_ASYNC_FWD_PACK=1 timeout 300 $B > $O/${TAG}_bench_asyncpack2.json 2>> $O/${TAG}_bench.err
for f in syncpack asyncpack syncpack2 asyncpack2; do python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_bench_$f.json")); print("%-16s %7.1f img/s  %6.2f ms  e2e %7.1f" % ("$f", d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$f: no result", e)
PY
done
timeout 300 python tools/timeline.py > $O/${TAG}_timeline.txt 2>&1; head -12 $O/${TAG}_timeline.txt
grep -v "OMP_NUM\|\*\*\*\*\|^$" $O/${TAG}_bench.err | tail -5 | cut -c1-300
