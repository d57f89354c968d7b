#!/bin/bash
# round 2, call u: load-batched 8-channel fold kernel, element-parallel unpack for few-tile tensors: tests that touch them,
# per-op table (fold was 1.21-1.24 ms, unpack 0.72-0.77 ms per step), bench
TAG=${1:-r02u}; O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_network.py tests/test_gpu_train_parity.py tests/test_gpu_guidance_grad.py -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
B="python bench.py --no-cpu-baseline --no-gpu-baseline --steps 20 --warmup 5"
timeout 300 $B > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 300 $B > $O/${TAG}_bench2.json 2>> $O/${TAG}_bench.err
for f in bench bench2; do python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_$f.json")); print("%-16s %7.1f img/s  %6.2f ms  e2e %7.1f" % ("$f", d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$f: no result", e)
PY
done
timeout 300 python tools/profile_ops.py > $O/${TAG}_profile_ops.log 2>&1; grep "fold \|unpack res64\|unpack upconv3\|FORWARD\|BACKWARD\|DTOD\|sum of\|by kind" $O/${TAG}_profile_ops.log
grep -v "OMP_NUM\|\*\*\*\*\|^$" $O/${TAG}_bench.err | tail -5 | cut -c1-300
