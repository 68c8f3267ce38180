#!/bin/bash
# generic same-box A/B of one environment switch: bash scripts/gpu_ab_env.sh TAG VAR "v0 v1 v0 v1" [pytest]
tag=$1; var=$2; arms=$3
out=gpurun_out/$tag
mkdir -p $out
if [ "$4" = "pytest" ]; then
  ( timeout 900 python -m pytest tests -m gpu -q ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
  grep -E "^FAILED|^ERROR" $out/pytest_gpu.log | head; tail -3 $out/pytest_gpu.log
fi
for arm in $arms; do
  ( env $var=$arm timeout 300 python bench.py --dump-ops $out/ops_$arm.json --no-gpu-reference --no-fast-ddpm --no-cpu-baseline --no-eval-kernels --no-e2e --no-strong ) > $out/bench_$arm.json 2> $out/bench_$arm.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/bench_$arm.json").read().strip().splitlines()[-1])
    print("== $var=$arm bench: ms_per_step %.3f  gemm %.3f" % (d["ms_per_step"], d["roofline"]["per_kernel_ms"]["pdr_gemm_fused"]))
except Exception as e:
    print("bench parse failed", e); print(open("$out/bench_$arm.err").read()[-600:])
PY
done
