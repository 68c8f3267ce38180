#!/bin/bash
# uniform warp index in the fused stage kernel: chain tests, then bench per-layer vs fused stages on one box
out=gpurun_out/${1:-r03d}
mkdir -p $out
( timeout 600 python -m pytest tests/test_chain_gpu.py tests/test_gemm_gpu.py -m gpu -q ) > $out/pytest_chain.log 2>&1; tail -3 $out/pytest_chain.log
for arm in 0 1 0 1; do
  ( PDR_STAGE_CHAIN=$arm timeout 300 python bench.py --dump-ops $out/ops_$arm.json --no-gpu-reference --no-fast-ddpm --no-cpu-baseline --no-eval-kernels --no-e2e --no-strong ) > $out/bench_$arm.json 2> $out/bench_$arm.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/bench_$arm.json").read().strip().splitlines()[-1])
    print("== PDR_STAGE_CHAIN=$arm bench: ms_per_step %.3f  %s" % (d["ms_per_step"], {k: round(v, 3) for k, v in list(d["roofline"]["per_kernel_ms"].items())[:4]}))
except Exception as e:
    print("bench parse failed", e); print(open("$out/bench_$arm.err").read()[-600:])
PY
done
