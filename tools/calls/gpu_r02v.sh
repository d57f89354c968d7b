#!/bin/bash
# round 2, call v: the round's last state -- whole GPU suite, smoke, default bench line (what the driver runs at round end)
TAG=${1:-r02v}; O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE-OK')" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$O/${TAG}_bench.json")); print("bench %7.1f img/s  %6.2f ms  e2e %7.1f  launches/step %d  clocks %s  lib %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"] / d["steps"], d["clocks"], d["gpu_library_baseline"]["bf16_channels_last"]))
PY
grep -v "OMP_NUM\|\*\*\*\*\|^$" $O/${TAG}_bench.err | tail -4 | cut -c1-300
