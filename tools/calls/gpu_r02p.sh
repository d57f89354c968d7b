#!/bin/bash
# round 2, call p: BatchNorm finalisation fused into the convolution tail (last CTA), unpack clears the wgrad scratch
# (46 fill launches gone), same Adam arithmetic in both launch modes: suite, determinism check, same-box A/B, smoke
TAG=${1:-r02p}; O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE-OK')" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
GDN_DETERMINISTIC=1 timeout 600 python tools/check_deterministic.py 10 4 > $O/${TAG}_deterministic.log 2>&1; tail -6 $O/${TAG}_deterministic.log
B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 20 --warmup 5"
GDN_FUSE_BNFIN=0 timeout 300 $B > $O/${TAG}_bench_nofin.json 2> $O/${TAG}_bench.err
timeout 300 $B > $O/${TAG}_bench_fin.json 2>> $O/${TAG}_bench.err
GDN_FUSE_BNFIN=0 timeout 300 $B > $O/${TAG}_bench_nofin2.json 2>> $O/${TAG}_bench.err
timeout 300 $B > $O/${TAG}_bench_fin2.json 2>> $O/${TAG}_bench.err
timeout 300 $B --workload train_dtod > $O/${TAG}_bench_train_dtod.json 2>> $O/${TAG}_bench.err
for f in nofin fin nofin2 fin2 train_dtod; do python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_bench_$f.json")); print("%-16s %7.1f img/s  %6.2f ms  e2e %7.1f  launches/step %d" % ("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"] / d["steps"]))
except Exception as e:
    print("$f: no result", e)
PY
done
timeout 300 python tools/profile_ops.py > $O/${TAG}_profile_ops.log 2>&1; grep "FORWARD\|BACKWARD\|DTOD\|sum of\|by kind" $O/${TAG}_profile_ops.log
grep -v "OMP_NUM\|\*\*\*\*\|^$" $O/${TAG}_bench.err | tail -5 | cut -c1-300
