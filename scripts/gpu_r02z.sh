#!/bin/bash
# lean gathered producer: parity (GEMM tests, then everything), probe timing, bench
out=gpurun_out/${1:-r02z}
mkdir -p $out
( timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x ) > $out/pytest_gemm.log 2>&1
rc=$?; tail -3 $out/pytest_gemm.log
if [ $rc -ne 0 ]; then grep -E "^E " $out/pytest_gemm.log | head -20; exit 0; fi
timeout 120 python scripts/prof_gather.py
( timeout 900 python -m pytest tests -m gpu -q ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
grep -E "^FAILED|^ERROR" $out/pytest_gpu.log | head; tail -3 $out/pytest_gpu.log
for arm in 0 1 0 1; do
  ( PDR_GEMM_TMA_GATHER=$arm timeout 300 python bench.py --dump-ops $out/ops_$arm.json --no-gpu-reference --no-fast-ddpm --no-cpu-baseline --no-eval-kernels --no-e2e --no-strong ) > $out/bench_$arm.json 2> $out/bench_$arm.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/bench_$arm.json").read().strip().splitlines()[-1])
    print("== TMA_GATHER=$arm (1 = old general loop + TMA, 0 = lean loop) bench: ms_per_step %.3f  gemm %.3f" % (d["ms_per_step"], d["roofline"]["per_kernel_ms"]["pdr_gemm_fused"]))
except Exception as e:
    print("bench parse failed", e); print(open("$out/bench_$arm.err").read()[-600:])
PY
done
