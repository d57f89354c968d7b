#!/bin/bash
# One gpurun call: GPU tests, per-op profile, bench (all workloads), ncu launch list, ncu full captures.
# Usage: tools/gpu_round.sh TAG
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.csv 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 300 python tools/profile_ops.py > $O/${TAG}_profile_ops.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err
timeout 300 python bench.py --workload infer --no-cpu-baseline --steps 20 --warmup 3 > $O/${TAG}_bench_infer.json 2>> $O/${TAG}_bench.err
timeout 300 python bench.py --workload train_dtod --no-cpu-baseline --steps 10 --warmup 3 > $O/${TAG}_bench_train_dtod.json 2>> $O/${TAG}_bench.err
timeout 300 python bench.py --workload infer_fullres --no-cpu-baseline --steps 10 --warmup 3 > $O/${TAG}_bench_infer_fullres.json 2>> $O/${TAG}_bench.err
# ncu: only the LAST step is profiled (cudaProfilerStart/Stop around it): construction, autotuning, warm-up are skipped
export GDN_GRAPH=0 GDN_PROFILE_LAST=1
NCU="ncu --profile-from-start off --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/${TAG}_launches.csv python tools/profile_step.py 3 > $O/${TAG}_ncu_launches.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:conv_igemm -c 12 -f -o $O/${TAG}_conv python tools/profile_step.py 3 > $O/${TAG}_ncu_conv.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:conv_wgrad -c 6 -f -o $O/${TAG}_wgrad python tools/profile_step.py 3 > $O/${TAG}_ncu_wgrad.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:'act_rows|act_up|bn_bwd|loss_kernel|adam_kernel|fold_rows|pack_tile|im2col' -c 24 -f -o $O/${TAG}_elem python tools/profile_step.py 3 > $O/${TAG}_ncu_elem.log 2>&1
unset GDN_PROFILE_LAST
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 3 -c 2 -f -o $O/${TAG}_conv64k9 python tools/profile_conv.py 20 > $O/${TAG}_ncu_conv64k9.log 2>&1
tail -3 $O/${TAG}_pytest.log; cat $O/${TAG}_bench.json $O/${TAG}_bench_reference.json $O/${TAG}_bench_infer.json $O/${TAG}_bench_train_dtod.json $O/${TAG}_bench_infer_fullres.json; tail -5 $O/${TAG}_bench.err
