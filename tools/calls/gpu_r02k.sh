#!/bin/bash
# round 2, call k: programmatic dependent launch (every kernel of the step chain) -- whole GPU suite with it on, then same-box A/B
TAG=${1:-r02k}; O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 20 --warmup 5"
GDN_PDL=0 timeout 300 $B > $O/${TAG}_bench_nopdl.json 2> $O/${TAG}_bench.err
GDN_PDL=1 timeout 300 $B > $O/${TAG}_bench_pdl.json 2>> $O/${TAG}_bench.err
GDN_PDL=0 timeout 300 $B > $O/${TAG}_bench_nopdl2.json 2>> $O/${TAG}_bench.err
GDN_PDL=1 timeout 300 $B > $O/${TAG}_bench_pdl2.json 2>> $O/${TAG}_bench.err
GDN_PDL=0 timeout 300 $B --workload infer > $O/${TAG}_bench_infer_nopdl.json 2>> $O/${TAG}_bench.err
GDN_PDL=1 timeout 300 $B --workload infer > $O/${TAG}_bench_infer_pdl.json 2>> $O/${TAG}_bench.err
GDN_PDL=1 timeout 300 python tools/profile_ops.py > $O/${TAG}_profile_ops.log 2>&1
for f in nopdl pdl nopdl2 pdl2 infer_nopdl infer_pdl; do python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_bench_$f.json")); print("%-14s %7.1f img/s  %6.2f ms  e2e %7.1f" % ("$f", d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$f: no result", e)
PY
done
grep -v "OMP_NUM\|\*\*\*\*\|^$" $O/${TAG}_bench.err | tail -5 | cut -c1-300
