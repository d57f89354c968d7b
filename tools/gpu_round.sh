#!/bin/bash
# One gpurun call: GPU tests, per-op profile, bench, ncu launch list, ncu full captures.  Usage: tools/gpu_round.sh TAG
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.csv 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 300 python tools/profile_ops.py > $O/${TAG}_profile_ops.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 300 python bench.py --workload infer --no-cpu-baseline --steps 20 --warmup 3 > $O/${TAG}_bench_infer.json 2>> $O/${TAG}_bench.err
GDN_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches.csv python tools/profile_step.py 2 > $O/${TAG}_ncu_launches.log 2>&1
GDN_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 1 -c 6 -f -o $O/${TAG}_conv python tools/profile_step.py 1 > $O/${TAG}_ncu_conv.log 2>&1
GDN_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 1 -c 3 -f -o $O/${TAG}_wgrad python tools/profile_step.py 1 > $O/${TAG}_ncu_wgrad.log 2>&1
GDN_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'act_forward|act_backward|bn_bwd_reduce|loss_kernel|adam_kernel|fold_grad' -s 40 -c 8 -f -o $O/${TAG}_elem python tools/profile_step.py 1 > $O/${TAG}_ncu_elem.log 2>&1
tail -3 $O/${TAG}_pytest.log; cat $O/${TAG}_bench.json; cat $O/${TAG}_bench_infer.json; tail -5 $O/${TAG}_bench.err
