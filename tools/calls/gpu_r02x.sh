#!/bin/bash
# round 2, call x (4 GPUs): scaling datapoint of the final state
TAG=${1:-r02x}; N=${2:-4}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench.err
python - <<PY
import json
d = json.load(open("$O/${TAG}_bench_n$N.json")); print("bench_n$N %8.1f img/s  %7.2f ms  e2e %8.1f  n_gpus %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["n_gpus"]))
PY
