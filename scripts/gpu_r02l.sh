#!/bin/bash
# A/B: first GEMM of wide stages cut into resident-weight pieces (PDR_SPLIT_WIDE) -- parity tests, then bench arms.
tag=${1:-r02l}
out=gpurun_out/$tag
mkdir -p $out
( timeout 600 python -m pytest tests/test_model_gpu.py tests/test_gemm_gpu.py tests/test_refinement_gpu.py -m gpu -q -x ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -3 $out/pytest_gpu.log
for arm in "PDR_SPLIT_WIDE=0" "PDR_SPLIT_WIDE=1" "PDR_SPLIT_WIDE=1 PDR_SPLIT_WIDE_MIN_ROWS=16384" $EXTRA_ARMS; do
  name=$(echo $arm | tr ' =' '__')
  ( env $arm timeout 300 python bench.py --dump-ops $out/ops_$name.json --no-gpu-reference --no-fast-ddpm --no-cpu-baseline --no-eval-kernels --no-e2e ) > $out/bench_$name.json 2> $out/bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/bench_$name.json").read().strip().splitlines()[-1])
    print("== $arm: ms_per_step %.3f  %s" % (d["ms_per_step"], {k: round(v, 3) for k, v in d["roofline"]["per_kernel_ms"].items()}))
    print("   gemm", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["roofline"]["tensor_core_kernels"]["pdr_gemm_fused"].items()})
except Exception as e:
    print("bench parse failed", e); print(open("$out/bench_$name.err").read()[-600:])
PY
done
