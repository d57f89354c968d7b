#!/bin/bash
# multi-GPU round: consistency check + bench at N (graph and eager).  Usage: tools/gpu_ddp.sh TAG N
TAG=${1:-r01}; N=${2:-2}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/check_ddp.py > $O/${TAG}_ddp_check.log 2>&1; echo "rc=$?" >> $O/${TAG}_ddp_check.log
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err; echo "rc=$?" >> $O/${TAG}_bench_n$N.err
GDN_GRAPH=0 timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > $O/${TAG}_bench_n${N}_eager.json 2> $O/${TAG}_bench_n${N}_eager.err
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
tail -5 $O/${TAG}_ddp_check.log; cat $O/${TAG}_bench_n$N.json; tail -3 $O/${TAG}_bench_n$N.err; cat $O/${TAG}_bench_n${N}_eager.json; cat $O/${TAG}_bench_n1.json
