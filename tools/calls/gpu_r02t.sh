#!/bin/bash
# round 2, call t (2 GPUs): final state on more than one GPU -- replica consistency at the bench batch (also in deterministic
# mode: graph == eager bit for bit across ranks), config-5 inference check, the -m gpu DDP test, bench at N = 1 and 2
TAG=${1:-r02t}; N=${2:-2}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
GDN_BATCH=20 timeout 400 $TR --master-port 29511 tools/check_ddp.py > $O/${TAG}_ddp_check.log 2>&1; echo "rc=$?" >> $O/${TAG}_ddp_check.log
GDN_DETERMINISTIC=1 GDN_BATCH=4 timeout 400 $TR --master-port 29516 tools/check_ddp.py > $O/${TAG}_ddp_check_det.log 2>&1; echo "rc=$?" >> $O/${TAG}_ddp_check_det.log
timeout 400 $TR --master-port 29514 tools/check_infer_ddp.py > $O/${TAG}_infer_ddp_check.log 2>&1; echo "rc=$?" >> $O/${TAG}_infer_ddp_check.log
timeout 600 python -m pytest tests/test_gpu_ddp.py -m gpu -q > $O/${TAG}_pytest_ddp.log 2>&1; tail -2 $O/${TAG}_pytest_ddp.log
timeout 300 python bench.py --no-cpu-baseline --no-gpu-baseline --steps 20 --warmup 5 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench.err
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_bench_n$N.json 2>> $O/${TAG}_bench.err
grep -h "DDP-OK\|teardown\|rc=\|Error" $O/${TAG}_ddp_check.log $O/${TAG}_ddp_check_det.log $O/${TAG}_infer_ddp_check.log | cut -c1-400
for f in bench_n1 bench_n$N; do python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_$f.json")); print("%-20s %8.1f img/s  %7.2f ms  e2e %8.1f  n_gpus %s" % ("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["n_gpus"]))
except Exception as e:
    print("$f: no result", e)
PY
done
grep -v "OMP_NUM\|\*\*\*\*\|^$\|barrier()\|return func" $O/${TAG}_bench.err | tail -5 | cut -c1-300
