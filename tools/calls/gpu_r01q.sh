#!/bin/bash
# Round-1 session-5 GPU call 1: new-feature tests first (all failures shown), the whole GPU suite, then benches.
TAG=${1:-r01q}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.csv 2>&1
timeout 600 python -m pytest tests/test_gpu_guidance_grad.py tests/test_gpu_demo.py -q -m gpu > $O/${TAG}_pytest_new.log 2>&1; echo "rc=$?" >> $O/${TAG}_pytest_new.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 300 python bench.py --workload train_guided --no-cpu-baseline --steps 10 --warmup 3 > $O/${TAG}_bench_train_guided.json 2>> $O/${TAG}_bench.err
timeout 300 python bench.py --workload demo --no-cpu-baseline --steps 50 --warmup 5 > $O/${TAG}_bench_demo.json 2>> $O/${TAG}_bench.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1
tail -30 $O/${TAG}_pytest_new.log; tail -5 $O/${TAG}_pytest.log; cat $O/${TAG}_bench.json $O/${TAG}_bench_train_guided.json $O/${TAG}_bench_demo.json; tail -5 $O/${TAG}_bench.err; tail -2 $O/${TAG}_smoke.log
