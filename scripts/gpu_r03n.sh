#!/bin/bash
# re-check two launch-time knobs after the codegen change: producer back-off and alternating-tile epilogue
out=gpurun_out/${1:-r03n}
mkdir -p $out
run() { # label env...
  label=$1; shift
  ( env "$@" timeout 300 python bench.py --no-gpu-reference --no-fast-ddpm --no-cpu-baseline --no-eval-kernels --no-e2e --no-strong ) > $out/bench_$label.json 2> $out/bench_$label.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/bench_$label.json").read().strip().splitlines()[-1])
    print("== $label: ms_per_step %.3f  gemm %.3f" % (d["ms_per_step"], d["roofline"]["per_kernel_ms"]["pdr_gemm_fused"]))
except Exception as e:
    print("bench parse failed", e)
PY
}
run base X=1
run sleep100 PDR_GEMM_PROD_SLEEP=100
run sleep300 PDR_GEMM_PROD_SLEEP=300
run epialt0 PDR_GEMM_EPI_ALT=0
run base2 X=1
run pdl1 PDR_PDL=1
