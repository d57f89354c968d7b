#!/bin/bash
# Round-1 session-5 GPU call 2: whole GPU suite, bench (+ reference arm, + A/B of the side-stream forward packs),
# timeline, ncu launch list and full captures of the final state.  Usage: tools/gpu_r01r.sh TAG
TAG=${1:-r01r}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.csv 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err
GDN_ASYNC_FWD_PACK=0 timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > $O/${TAG}_bench_syncpack.json 2>> $O/${TAG}_bench.err
timeout 300 python tools/timeline.py > $O/${TAG}_timeline.txt 2>&1
export GDN_GRAPH=0 GDN_PROFILE_LAST=1
NCU="ncu --profile-from-start off --clock-control none"
timeout 500 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/${TAG}_launches.csv python tools/profile_step.py 3 > $O/${TAG}_ncu_launches.log 2>&1
timeout 500 $NCU --set full --import-source on -k regex:conv_igemm -c 12 -f -o $O/${TAG}_conv python tools/profile_step.py 3 > $O/${TAG}_ncu_conv.log 2>&1
timeout 500 $NCU --set full --import-source on -k regex:conv_wgrad -c 6 -f -o $O/${TAG}_wgrad python tools/profile_step.py 3 > $O/${TAG}_ncu_wgrad.log 2>&1
timeout 500 $NCU --set full --import-source on -k regex:'act_rows|act_up|bn_bwd|loss_kernel|adam_kernel|fold_rows|pack_t|im2col|unpack|eigen' -c 28 -f -o $O/${TAG}_elem python tools/profile_step.py 3 > $O/${TAG}_ncu_elem.log 2>&1
unset GDN_PROFILE_LAST
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'resize_|bytescale|minmax_k|frozen_bwd|fold_thin|sqdiff_grad|tanh_chain|preprocess_u8' -c 40 -f -o $O/${TAG}_extras python tools/profile_extras.py > $O/${TAG}_ncu_extras.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 3 -c 2 -f -o $O/${TAG}_conv64k9 python tools/profile_conv.py 20 > $O/${TAG}_ncu_conv64k9.log 2>&1
tail -3 $O/${TAG}_pytest.log; cat $O/${TAG}_bench.json $O/${TAG}_bench_reference.json $O/${TAG}_bench_syncpack.json; tail -5 $O/${TAG}_bench.err; head -1 $O/${TAG}_timeline.txt; ls -la $O | tail -20
