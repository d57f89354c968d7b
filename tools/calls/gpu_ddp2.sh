#!/bin/bash
# 2-GPU check of the final state: replica consistency (tools/check_ddp.py) + the driver's N=2 bench launch
TAG=${1:-r01u}; N=${2:-2}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/check_ddp.py > $O/${TAG}_ddp_check.log 2>&1; echo "rc=$?" >> $O/${TAG}_ddp_check.log
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err; echo "rc=$?" >> $O/${TAG}_bench_n$N.err
timeout 120 $TR --master-port 29513 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > $O/${TAG}_bench_ref_n$N.json 2>> $O/${TAG}_bench_n$N.err
tail -4 $O/${TAG}_ddp_check.log | cut -c1-600; cat $O/${TAG}_bench_n$N.json; tail -3 $O/${TAG}_bench_n$N.err; cat $O/${TAG}_bench_ref_n$N.json
