#!/bin/bash
# round 2, call i (8 GPUs): config 5 on 8 GPUs, replica consistency at the bench batch, scaling bench + NCCL CTA cap A/B
TAG=${1:-r02i}; N=8; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29514 tools/check_infer_ddp.py > $O/${TAG}_infer_ddp_check.log 2>&1; echo "rc=$?" >> $O/${TAG}_infer_ddp_check.log
GDN_BATCH=20 timeout 400 $TR --master-port 29511 tools/check_ddp.py > $O/${TAG}_ddp_check.log 2>&1; echo "rc=$?" >> $O/${TAG}_ddp_check.log
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_bench_n8.json 2> $O/${TAG}_bench.err
NCCL_MAX_CTAS=4 timeout 300 $TR --master-port 29516 bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_bench_n8_maxctas4.json 2>> $O/${TAG}_bench.err
NCCL_MAX_CTAS=16 timeout 300 $TR --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_bench_n8_maxctas16.json 2>> $O/${TAG}_bench.err
timeout 300 $TR --master-port 29515 bench.py --gpus $N --workload infer_fullres --steps 10 --warmup 3 > $O/${TAG}_bench_fullres_n8.json 2>> $O/${TAG}_bench.err
timeout 300 python bench.py --no-cpu-baseline --no-gpu-baseline --steps 20 --warmup 5 > $O/${TAG}_bench_n1.json 2>> $O/${TAG}_bench.err
grep -h "DDP-OK\|teardown\|rc=\|Error" $O/${TAG}_ddp_check.log $O/${TAG}_infer_ddp_check.log | cut -c1-400
for f in bench_n1 bench_n8 bench_n8_maxctas4 bench_n8_maxctas16 bench_fullres_n8; do python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_$f.json")); print("%-20s %8.1f img/s  %7.2f ms  e2e %8.1f  n_gpus %s" % ("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["n_gpus"]))
except Exception as e:
    print("$f: no result", e)
PY
done
grep -v "OMP_NUM\|\*\*\*\*\|^$" $O/${TAG}_bench.err | tail -6 | cut -c1-300
