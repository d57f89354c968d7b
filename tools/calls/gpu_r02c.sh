#!/bin/bash
# round 2, call c: fused BN-backward statistics + bf16 folds on hardware: suite, bench A/B, per-op table, trajectory probe
TAG=${1:-r02c}; O=gpurun_out; mkdir -p $O
B="python bench.py --no-cpu-baseline --steps 10 --warmup 3"
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
timeout 300 $B > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench.err
GDN_FUSE_BNBWD=0 timeout 300 $B > $O/${TAG}_bench_nofuse.json 2>> $O/${TAG}_bench.err
GDN_FOLD_BF16=0 timeout 300 $B > $O/${TAG}_bench_fold32.json 2>> $O/${TAG}_bench.err
timeout 300 $B > $O/${TAG}_bench_default2.json 2>> $O/${TAG}_bench.err
timeout 300 python tools/profile_ops.py > $O/${TAG}_profile_ops.log 2>&1
timeout 600 python -m tests.parity_probe traj --b 4 --steps 50 --kinds init > $O/${TAG}_parity_traj.log 2>&1
timeout 400 python -m tests.parity_probe grad --b 4 --kinds warm > $O/${TAG}_parity_grad.log 2>&1
for f in default nofuse fold32 default2; do python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_bench_$f.json")); print("%-10s %7.1f img/s  %6.2f ms  e2e %7.1f" % ("$f", d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$f: no result", e)
PY
done
grep -E "by kind|====" $O/${TAG}_profile_ops.log; tail -4 $O/${TAG}_parity_traj.log | cut -c1-400; grep "whole\|tensors with" $O/${TAG}_parity_grad.log | cut -c1-400
