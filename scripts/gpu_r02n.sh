#!/bin/bash
# A/B: producers of raw GEMM operands round to TF32 (PDR_ROUND_TABLES) -- error distribution of the TF32 step, then the full suite + bench.
tag=${1:-r02n}
out=gpurun_out/$tag
mkdir -p $out
for r in 0 1; do
  ( PDR_ROUND_TABLES=$r timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -s -k "tf32 or benchmarked" ) > $out/pytest_round_$r.log 2>&1
  echo "== PDR_ROUND_TABLES=$r"; grep -E "TF32|cd_t\(|passed|failed" $out/pytest_round_$r.log
done
( timeout 900 python -m pytest tests -m gpu -q ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -3 $out/pytest_gpu.log
( timeout 300 python bench.py --dump-ops $out/ops.json --no-gpu-reference --no-fast-ddpm --no-cpu-baseline --no-eval-kernels --no-e2e ) > $out/bench.json 2> $out/bench.err
python - <<PY
import json
try:
    d = json.loads(open("$out/bench.json").read().strip().splitlines()[-1])
    print("== bench: ms_per_step %.3f  %s" % (d["ms_per_step"], {k: round(v, 3) for k, v in d["roofline"]["per_kernel_ms"].items()}))
except Exception as e:
    print("bench parse failed", e); print(open("$out/bench.err").read()[-600:])
PY
