#!/bin/bash
# round 2, call e: evidence -- true per-kernel device times (ncu launch list of one eager step), device-scheduled
# timeline, ncu --set full of the shipped dominant conv alone and of the HBM-side kernels at bench shape, metric/loss kernels
TAG=${1:-r02e}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_train_parity.py -x -q > $O/${TAG}_pytest_parity.log 2>&1; tail -3 $O/${TAG}_pytest_parity.log
timeout 300 python tools/timeline.py > $O/${TAG}_timeline.txt 2>&1; head -1 $O/${TAG}_timeline.txt
timeout 200 python tools/profile_metrics.py > $O/${TAG}_profile_metrics.log 2>&1; cat $O/${TAG}_profile_metrics.log
export GDN_GRAPH=0 GDN_PROFILE_LAST=1
NCU="ncu --profile-from-start off --clock-control none"
timeout 500 $NCU --metrics gpu__time_duration.sum,launch__grid_size --csv --log-file $O/${TAG}_launches.csv python tools/profile_step.py 3 > $O/${TAG}_ncu_launches.log 2>&1
python tools/launch_summary.py $O/${TAG}_launches.csv > $O/${TAG}_launches_by_kernel.txt 2>&1; head -30 $O/${TAG}_launches_by_kernel.txt
timeout 600 $NCU --set full --import-source on -k regex:'bn_bwd|fold_rows|act_up_rows|act_rows|adam_kernel|loss_kernel|absdiff|unpack_tile|pack_table|im2col|splitk_combine' -c 60 -f -o /tmp/${TAG}_elem python tools/profile_step.py 3 > $O/${TAG}_ncu_elem.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_elem.ncu-rep > $O/${TAG}_ncu_elem.summary.txt 2>&1
unset GDN_PROFILE_LAST
timeout 300 $NCU --set full --import-source on -k regex:conv_igemm -c 2 -f -o $O/${TAG}_conv64k9 python tools/profile_conv.py 20 > $O/${TAG}_ncu_conv64k9.log 2>&1
python tools/ncu_summary.py $O/${TAG}_conv64k9.ncu-rep --traffic-json $O/${TAG}_dominant_conv_traffic.json conv_igemm > $O/${TAG}_ncu_conv64k9.summary.txt 2>&1
cat $O/${TAG}_ncu_conv64k9.summary.txt
timeout 300 $NCU --set full --import-source on -k regex:'eigen_metrics|loss_kernel|absdiff' -c 6 -f -o /tmp/${TAG}_metrics python tools/profile_metrics.py > $O/${TAG}_ncu_metrics.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_metrics.ncu-rep > $O/${TAG}_ncu_metrics.summary.txt 2>&1
cat $O/${TAG}_ncu_metrics.summary.txt; du -sh $O
