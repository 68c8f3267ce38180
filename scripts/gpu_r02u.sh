#!/bin/bash
# packed-pair statistics in the TMA-store epilogue: parity, micro timing, bench
out=gpurun_out/${1:-r02u}
mkdir -p $out
( timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x ) > $out/pytest_gemm.log 2>&1
rc=$?; tail -3 $out/pytest_gemm.log
if [ $rc -ne 0 ]; then grep -E "^E " $out/pytest_gemm.log | head -20; fi
timeout 120 python scripts/prof_gather.py
( timeout 900 python -m pytest tests -m gpu -q ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
grep -E "^FAILED|^ERROR" $out/pytest_gpu.log | head; tail -3 $out/pytest_gpu.log
for arm in 1 1; do
  ( timeout 300 python bench.py --dump-ops $out/ops_$arm.json --no-gpu-reference --no-fast-ddpm --no-cpu-baseline --no-eval-kernels --no-e2e --no-strong ) > $out/bench_$arm.json 2> $out/bench_$arm.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/bench_$arm.json").read().strip().splitlines()[-1])
    print("== bench: ms_per_step %.3f  %s" % (d["ms_per_step"], {k: round(v, 3) for k, v in d["roofline"]["per_kernel_ms"].items()}))
except Exception as e:
    print("bench parse failed", e); print(open("$out/bench_$arm.err").read()[-600:])
PY
done
