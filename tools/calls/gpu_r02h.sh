#!/bin/bash
# round 2, call h (N GPUs): replica consistency at the bench batch, config-5 inference check, bench at N, clean teardown
TAG=${1:-r02h}; N=${2:-2}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
GDN_BATCH=20 timeout 400 $TR --master-port 29511 tools/check_ddp.py > $O/${TAG}_ddp_check.log 2>&1; echo "rc=$?" >> $O/${TAG}_ddp_check.log
timeout 400 $TR --master-port 29514 tools/check_infer_ddp.py > $O/${TAG}_infer_ddp_check.log 2>&1; echo "rc=$?" >> $O/${TAG}_infer_ddp_check.log
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err; echo "rc=$?" >> $O/${TAG}_bench_n$N.err
timeout 300 $TR --master-port 29515 bench.py --gpus $N --workload infer_fullres --steps 10 --warmup 3 > $O/${TAG}_bench_fullres_n$N.json 2>> $O/${TAG}_bench_n$N.err; echo "rc=$?" >> $O/${TAG}_bench_n$N.err
timeout 200 $TR --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/${TAG}_bench_ref_n$N.json 2>> $O/${TAG}_bench_n$N.err
grep -h "DDP-OK\|teardown\|rc=\|Error" $O/${TAG}_ddp_check.log $O/${TAG}_infer_ddp_check.log | cut -c1-400
for f in bench_n$N bench_fullres_n$N bench_ref_n$N; do python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_$f.json")); print("%-20s %8.1f img/s  %7.2f ms  e2e %8.1f  n_gpus %s" % ("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["n_gpus"]))
except Exception as e:
    print("$f: no result", e)
PY
done
tail -6 $O/${TAG}_bench_n$N.err | cut -c1-300
