#!/bin/bash
# round 2, call m: second epilogue warp group (algo bit 28) -- kernel tests, variant sweep, same-box A/B; cp.async head gather
TAG=${1:-r02m}; O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 20 --warmup 5"
GDN_EW8=0 timeout 300 $B > $O/${TAG}_bench_ew4.json 2> $O/${TAG}_bench.err
GDN_EW8=1 timeout 300 $B > $O/${TAG}_bench_ew8.json 2>> $O/${TAG}_bench.err
GDN_EW8=0 timeout 300 $B > $O/${TAG}_bench_ew4b.json 2>> $O/${TAG}_bench.err
GDN_EW8=1 timeout 300 $B > $O/${TAG}_bench_ew8b.json 2>> $O/${TAG}_bench.err
GDN_EW8=0 timeout 300 $B --workload infer > $O/${TAG}_bench_infer_ew4.json 2>> $O/${TAG}_bench.err
GDN_EW8=1 timeout 300 $B --workload infer > $O/${TAG}_bench_infer_ew8.json 2>> $O/${TAG}_bench.err
GDN_EW8=1 timeout 300 $B --workload train_dtod > $O/${TAG}_bench_train_dtod_ew8.json 2>> $O/${TAG}_bench.err
for f in ew4 ew8 ew4b ew8b infer_ew4 infer_ew8 train_dtod_ew8; do python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_bench_$f.json")); print("%-16s %7.1f img/s  %6.2f ms  e2e %7.1f" % ("$f", d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$f: no result", e)
PY
done
timeout 300 python tools/profile_ops.py > $O/${TAG}_profile_ops.log 2>&1; grep "upconv4\|head\|FORWARD\|BACKWARD\|DTOD\|sum of\|by kind" $O/${TAG}_profile_ops.log
timeout 400 python tools/sweep_conv.py > $O/${TAG}_sweep_conv.log 2>&1; grep -A3 "^==" $O/${TAG}_sweep_conv.log | cut -c1-100
grep -v "OMP_NUM\|\*\*\*\*\|^$" $O/${TAG}_bench.err | tail -5 | cut -c1-300
