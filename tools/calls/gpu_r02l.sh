#!/bin/bash
# round 2, call l: 64->1 heads as taps-as-N 1x1 conv + shifted sum (tests, same-box A/B, per-op table), reproducible oracle
# warm-up (parity tests), and ncu --set full (source counters) of two short-reduction convolutions
TAG=${1:-r02l}; O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log; grep -h "train fwd" $O/${TAG}_pytest.log
B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 20 --warmup 5"
GDN_HEAD_TAPS=0 timeout 300 $B > $O/${TAG}_bench_headconv.json 2> $O/${TAG}_bench.err
GDN_HEAD_TAPS=1 timeout 300 $B > $O/${TAG}_bench_headtaps.json 2>> $O/${TAG}_bench.err
GDN_HEAD_TAPS=0 timeout 300 $B --workload infer > $O/${TAG}_bench_infer_headconv.json 2>> $O/${TAG}_bench.err
GDN_HEAD_TAPS=1 timeout 300 $B --workload infer > $O/${TAG}_bench_infer_headtaps.json 2>> $O/${TAG}_bench.err
timeout 300 python tools/profile_ops.py > $O/${TAG}_profile_ops.log 2>&1; grep "upconv4\|head\|FORWARD\|BACKWARD\|sum of" $O/${TAG}_profile_ops.log
for f in headconv headtaps infer_headconv infer_headtaps; do python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_bench_$f.json")); print("%-16s %7.1f img/s  %6.2f ms  e2e %7.1f" % ("$f", d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$f: no result", e)
PY
done
NCU="ncu --clock-control none --set full --import-source on -k regex:conv_igemm -s 4 -c 1 -f"
timeout 300 $NCU -o $O/${TAG}_conv1x1_concat python tools/probe_conv_case.py "128->64 k1" 0x1010001 > $O/${TAG}_ncu_conv1x1.log 2>&1; tail -2 $O/${TAG}_ncu_conv1x1.log
timeout 300 $NCU -o $O/${TAG}_conv512_8x26 python tools/probe_conv_case.py "512->512 k3 8x26 train" 0x1040001 > $O/${TAG}_ncu_conv512.log 2>&1; tail -2 $O/${TAG}_ncu_conv512.log
grep -v "OMP_NUM\|\*\*\*\*\|^$" $O/${TAG}_bench.err | tail -5 | cut -c1-300; du -sh $O
