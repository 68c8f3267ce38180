#!/bin/bash
# light last check of the tree as it will be judged: GEMM / model tests, smoke, warm-step bench
out=gpurun_out/${1:-last}
mkdir -p $out
( timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py -m gpu -q ) > $out/pytest.log 2>&1; tail -2 $out/pytest.log
( timeout 200 python __graft_entry__.py smoke ) > $out/smoke.log 2>&1; tail -1 $out/smoke.log | cut -c1-120
( timeout 300 python bench.py --no-gpu-reference --no-fast-ddpm --no-cpu-baseline --no-eval-kernels --no-strong ) > $out/bench.json 2> $out/bench.err
python - <<PY
import json
d = json.loads(open("$out/bench.json").read().strip().splitlines()[-1])
print("== bench: ms_per_step %.3f value %.3f e2e %.3f frac %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"]))
PY
