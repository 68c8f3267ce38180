#!/bin/bash
# A/B: programmatic dependent launch (PDR_PDL) of the GEMM + the small row kernels -- full suite with it on, then bench off/on/off/on.
tag=${1:-r02o}
out=gpurun_out/$tag
mkdir -p $out
( PDR_PDL=4 timeout 900 python -m pytest tests -m gpu -q -x ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
grep -E "^FAILED|^ERROR" $out/pytest_gpu.log | head; tail -3 $out/pytest_gpu.log
for arm in 2 4 2 4; do
  ( PDR_PDL=$arm timeout 300 python bench.py --dump-ops $out/ops_$arm.json --no-gpu-reference --no-fast-ddpm --no-cpu-baseline --no-eval-kernels --no-e2e --no-strong ) > $out/bench_$arm.json 2> $out/bench_$arm.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/bench_$arm.json").read().strip().splitlines()[-1])
    print("== PDR_PDL=$arm bench: ms_per_step %.3f  %s" % (d["ms_per_step"], {k: round(v, 3) for k, v in d["roofline"]["per_kernel_ms"].items()}))
except Exception as e:
    print("bench parse failed", e); print(open("$out/bench_$arm.err").read()[-600:])
PY
done
