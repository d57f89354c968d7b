#!/bin/bash
# final-state round: GPU suite, bench + reference arm, ncu launch list + full captures SUMMARISED ON THE BOX
# (gpurun_out/ is limited to 64 MiB: the .ncu-rep files are deleted after tools/ncu_summary.py has read them,
# except the small single-kernel capture).  Usage: tools/gpu_r01t.sh TAG
TAG=${1:-r01t}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.csv 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err
export GDN_GRAPH=0 GDN_PROFILE_LAST=1
NCU="ncu --profile-from-start off --clock-control none"
timeout 400 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/${TAG}_launches.csv python tools/profile_step.py 3 > $O/${TAG}_ncu_launches.log 2>&1
python tools/launch_summary.py $O/${TAG}_launches.csv > $O/${TAG}_launches_by_kernel.txt 2>&1
timeout 400 $NCU --set full --import-source on -k regex:conv_igemm -c 12 -f -o /tmp/${TAG}_conv python tools/profile_step.py 3 > $O/${TAG}_ncu_conv.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_conv.ncu-rep > $O/${TAG}_ncu_conv.summary.txt 2>&1
unset GDN_PROFILE_LAST
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'resize_|bytescale|minmax_k|frozen_bwd|fold_thin|sqdiff_grad|tanh_chain|preprocess_u8' -c 40 -f -o /tmp/${TAG}_extras python tools/profile_extras.py > $O/${TAG}_ncu_extras.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_extras.ncu-rep > $O/${TAG}_ncu_extras.summary.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 3 -c 2 -f -o $O/${TAG}_conv64k9 python tools/profile_conv.py 20 > $O/${TAG}_ncu_conv64k9.log 2>&1
python tools/ncu_summary.py $O/${TAG}_conv64k9.ncu-rep > $O/${TAG}_ncu_conv64k9.summary.txt 2>&1
rm -f $O/r01r_*.ncu-rep
tail -3 $O/${TAG}_pytest.log; cat $O/${TAG}_bench.json $O/${TAG}_bench_reference.json; tail -3 $O/${TAG}_bench.err; head -12 $O/${TAG}_launches_by_kernel.txt; cat $O/${TAG}_ncu_conv64k9.summary.txt; du -sh $O
