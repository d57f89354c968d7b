#!/bin/bash
# FIRST GPU call of round 2 (prepared at the end of round 1, when the GPU budget was spent): settle the open A/Bs on
# ONE box, back to back, and produce the per-op TFLOP/s table.  ~8 min.  Usage: tools/gpu_r02a.sh [TAG]
TAG=${1:-r02a}; O=gpurun_out; mkdir -p $O
B="python bench.py --no-cpu-baseline --steps 10 --warmup 3"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.csv 2>&1
# 1. default state: whole suite + bench
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 300 $B > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench.err
# 2. first versions of the fold / re-pack kernels (round-1 A/B that was left inside box-to-box noise)
GDN_FOLD_V1=1 GDN_PACK_V1=1 timeout 300 $B > $O/${TAG}_bench_v1kernels.json 2>> $O/${TAG}_bench.err
GDN_FOLD_V1=1 timeout 300 $B > $O/${TAG}_bench_foldv1.json 2>> $O/${TAG}_bench.err
# 3. transposed fp32 epilogue (never run on hardware before): correctness first, then the bench
GDN_EPI_T=1 timeout 600 python -m pytest tests/test_gpu_new_kernels.py tests/test_gpu_kernels.py tests/test_gpu_network.py tests/test_gpu_guidance_grad.py -q -m gpu > $O/${TAG}_pytest_epi_t.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_epi_t.log
GDN_EPI_T=1 timeout 300 $B > $O/${TAG}_bench_epi_t.json 2>> $O/${TAG}_bench.err
timeout 300 $B > $O/${TAG}_bench_default2.json 2>> $O/${TAG}_bench.err        # default again: drift of the box during the call
# 4. per-op table (time, GFLOP, TFLOP/s, M tiles) and the library GPU baseline
timeout 300 python tools/profile_ops.py > $O/${TAG}_profile_ops.log 2>&1
GDN_EPI_T=1 timeout 300 python tools/profile_ops.py > $O/${TAG}_profile_ops_epi_t.log 2>&1
timeout 400 python tools/bench_torch_gpu.py > $O/${TAG}_bench_torch.jsonl 2> $O/${TAG}_bench_torch.err
tail -3 $O/${TAG}_pytest.log; tail -6 $O/${TAG}_pytest_epi_t.log
for f in default v1kernels foldv1 epi_t default2; do python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_bench_$f.json")); print("%-10s %7.1f img/s  %6.2f ms  e2e %7.1f" % ("$f", d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$f: no result", e)
PY
done
grep -E "tensor-core launches|TFLOP/s:" $O/${TAG}_profile_ops.log | head -20; cat $O/${TAG}_bench_torch.jsonl; du -sh $O
