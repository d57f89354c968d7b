#!/bin/bash
# ncu evidence of one eager training step (launch list + full captures), summarised ON the box so that only small
# text files travel back (gpurun_out is capped at 64 MiB).  Usage: tools/gpu_profile.sh TAG
TAG=${1:-r01}
O=gpurun_out; T=/tmp/ncu_$TAG
mkdir -p $O $T
export GDN_GRAPH=0 GDN_PROFILE_LAST=1
NCU="ncu --profile-from-start off --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/${TAG}_launches.csv python tools/profile_step.py 3 > $O/${TAG}_ncu_launches.log 2>&1
python tools/launch_summary.py $O/${TAG}_launches.csv > $O/${TAG}_launches_by_kernel.txt 2>&1
timeout 600 $NCU --set full --import-source on -k regex:conv_igemm -c 10 -f -o $T/conv python tools/profile_step.py 3 > $O/${TAG}_ncu_conv.log 2>&1
timeout 600 $NCU --set full -k regex:conv_wgrad -c 5 -f -o $T/wgrad python tools/profile_step.py 3 > $O/${TAG}_ncu_wgrad.log 2>&1
timeout 600 $NCU --set full -k regex:'act_rows|act_up|bn_bwd|loss_kernel|adam_kernel|fold_rows|pack_table|im2col|unpack_tile' -c 24 -f -o $T/elem python tools/profile_step.py 3 > $O/${TAG}_ncu_elem.log 2>&1
unset GDN_PROFILE_LAST
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 30 -c 2 -f -o $T/conv64k9 python tools/profile_conv.py 20 > $O/${TAG}_ncu_conv64k9.log 2>&1
for k in conv wgrad elem conv64k9; do python tools/ncu_summary.py $T/$k.ncu-rep > $O/${TAG}_ncu_$k.summary.txt 2>&1; done
cp $T/conv64k9.ncu-rep $O/${TAG}_conv64k9.ncu-rep 2>/dev/null
ls -la $T; du -sh $O
