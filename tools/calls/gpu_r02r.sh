#!/bin/bash
# round 2, call r: FINAL-STATE evidence -- bench lines of every workload (default line with the library-baseline and CPU
# legs), ncu launch list of one eager step, ncu --set full of the shipped dominant conv alone (roofline.traffic), of the
# metric / loss kernels and of the new head-gather + short-reduction launches inside the step, device timeline
TAG=${1:-r02r}; O=gpurun_out; mkdir -p $O
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
for wl in infer infer_fullres train_dtod train_guided demo; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-gpu-baseline --steps 20 --warmup 5 > $O/${TAG}_bench_$wl.json 2>> $O/${TAG}_bench.err
done
for f in bench bench_infer bench_infer_fullres bench_train_dtod bench_train_guided bench_demo; do python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_$f.json")); print("%-22s %8.1f %s  %7.2f ms  e2e %8.1f  launches/step %d" % ("$f", d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"] / d["steps"]))
except Exception as e:
    print("$f: no result", e)
PY
done
timeout 300 python tools/timeline.py > $O/${TAG}_timeline.txt 2>&1; head -1 $O/${TAG}_timeline.txt
timeout 200 python tools/profile_metrics.py > $O/${TAG}_profile_metrics.log 2>&1; cat $O/${TAG}_profile_metrics.log
export GDN_GRAPH=0 GDN_PROFILE_LAST=1
NCU="ncu --profile-from-start off --clock-control none"
timeout 500 $NCU --metrics gpu__time_duration.sum,launch__grid_size --csv --log-file $O/${TAG}_launches.csv python tools/profile_step.py 3 > $O/${TAG}_ncu_launches.log 2>&1
python tools/launch_summary.py $O/${TAG}_launches.csv > $O/${TAG}_launches_by_kernel.txt 2>&1; head -34 $O/${TAG}_launches_by_kernel.txt
timeout 600 $NCU --set full --import-source on -k regex:'head_gather|bn_bwd|fold_rows|act_up_rows|up2x|act_rows|adam_kernel|unpack_tile|pack_table|im2col' -c 220 -f -o /tmp/${TAG}_elem python tools/profile_step.py 3 > $O/${TAG}_ncu_elem.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_elem.ncu-rep > $O/${TAG}_ncu_elem.summary.txt 2>&1
unset GDN_PROFILE_LAST
timeout 300 $NCU --set full --import-source on -k regex:conv_igemm -c 2 -f -o $O/${TAG}_conv64k9 python tools/profile_conv.py 20 > $O/${TAG}_ncu_conv64k9.log 2>&1
python tools/ncu_summary.py $O/${TAG}_conv64k9.ncu-rep --traffic-json $O/${TAG}_dominant_conv_traffic.json conv_igemm > $O/${TAG}_ncu_conv64k9.summary.txt 2>&1
cat $O/${TAG}_ncu_conv64k9.summary.txt
timeout 300 $NCU --set full --import-source on -k regex:'metrics_|loss_|absdiff' -c 16 -f -o /tmp/${TAG}_metrics python tools/profile_metrics.py > $O/${TAG}_ncu_metrics.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_metrics.ncu-rep > $O/${TAG}_ncu_metrics.summary.txt 2>&1
cut -c1-260 $O/${TAG}_ncu_metrics.summary.txt
grep -v "OMP_NUM\|\*\*\*\*\|^$" $O/${TAG}_bench.err | tail -4 | cut -c1-300; du -sh $O
