#!/bin/bash
# A/B: TMA tile::gather4 for the gathered table chunks (PDR_GEMM_TMA_GATHER) on top of the late-trigger PDL default.
tag=${1:-r02p}
out=gpurun_out/$tag
mkdir -p $out
( timeout 180 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k gathered ) > $out/pytest_gathered.log 2>&1
rc=$?; tail -5 $out/pytest_gathered.log
if [ $rc -ne 0 ]; then echo "gathered GEMM test failed (rc $rc): stopping"; grep -E "^E " $out/pytest_gathered.log | head -20; exit 0; fi
( timeout 900 python -m pytest tests -m gpu -q -x ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
grep -E "^FAILED|^ERROR" $out/pytest_gpu.log | head; tail -3 $out/pytest_gpu.log
for arm in 0 1 0 1; do
  ( PDR_GEMM_TMA_GATHER=$arm timeout 300 python bench.py --dump-ops $out/ops_$arm.json --no-gpu-reference --no-fast-ddpm --no-cpu-baseline --no-eval-kernels --no-e2e --no-strong ) > $out/bench_$arm.json 2> $out/bench_$arm.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/bench_$arm.json").read().strip().splitlines()[-1])
    print("== PDR_GEMM_TMA_GATHER=$arm bench: ms_per_step %.3f  %s" % (d["ms_per_step"], {k: round(v, 3) for k, v in d["roofline"]["per_kernel_ms"].items()}))
except Exception as e:
    print("bench parse failed", e); print(open("$out/bench_$arm.err").read()[-600:])
PY
done
