#!/bin/bash
# zero-code re-check of two engine switches after the codegen change: pooling inside the score GEMM's epilogue, residual fold
out=gpurun_out/${1:-r03o}
mkdir -p $out
run() { label=$1; shift
  ( env "$@" timeout 300 python bench.py --no-gpu-reference --no-fast-ddpm --no-cpu-baseline --no-eval-kernels --no-e2e --no-strong ) > $out/bench_$label.json 2> $out/bench_$label.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/bench_$label.json").read().strip().splitlines()[-1])
    print("== $label: ms_per_step %.3f  gemm %.3f pool %.3f" % (d["ms_per_step"], d["roofline"]["per_kernel_ms"]["pdr_gemm_fused"], d["roofline"]["per_kernel_ms"].get("pdr_attention_pool", 0)))
except Exception as e:
    print("$label: bench parse failed", e); print(open("$out/bench_$label.err").read()[-400:])
PY
}
run base X=1
run fusepool PDR_FUSE_POOL=1
run nofold PDR_FOLD_RES=0
