#!/bin/bash
# last visit of the round: the committed tree -- full GPU suite, smoke, default bench line
out=gpurun_out/${1:-final_check}
mkdir -p $out
( timeout 900 python -m pytest tests -m gpu -q ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log; tail -3 $out/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke ) > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log; tail -2 $out/smoke.log
( timeout 900 python bench.py ) > $out/bench.json 2> $out/bench.err; echo "bench exit $?"
python - <<PY
import json
d = json.loads(open("$out/bench.json").read().strip().splitlines()[-1])
print("== bench: ms_per_step %.3f value %.3f e2e %.3f frac %.3f launches %s clocks %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"]))
PY
