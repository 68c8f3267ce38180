#!/bin/bash
# First GPU visit of the next round: A/B of the four opt-in experiments prepared at the end of round 1 (none of them has
# run on a GPU yet -- each gets the GEMM / model parity tests under its switch first, under a hard timeout, and its
# bench arm only if they pass).  ~5 minutes of box time.
#   PDR_GEMM_IDX_RING=1   gathered-A indices through a per-warp shared-memory ring (gemm_tc.cuh, GRING)
#   PDR_GEMM_TAIL_X=1     raw K-tail chunks copied by the transform warps (gemm_tc.cuh, TAILX)
#   PDR_GEOM_OVERLAP=1    geometry chain on a side stream next to the first mapper block (fused.py, PdrGemmArgs.max_ctas)
#   PDR_GEMM_GN_FUSED=1   pdr_gn_finalize folded into the GEMM that produces its last source (gemm_tc.cuh GNF, PdrGemmArgs.gn_fused)
tag=${1:-r02ab}
out=gpurun_out/$tag
mkdir -p $out
arms=("-")
for sw in PDR_GEMM_IDX_RING=1 PDR_GEMM_TAIL_X=1 PDR_GEOM_OVERLAP=1 PDR_GEMM_GN_FUSED=1; do
  name=${sw%%=*}
  ( env $sw timeout 240 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py tests/test_refinement_gpu.py -m gpu -x -q --deselect tests/test_model_gpu.py::test_tf32_chain_lands_where_the_fp32_chain_does ) \
      > $out/pytest_$name.log 2>&1
  rc=$?
  echo "$sw pytest exit $rc: $(tail -1 $out/pytest_$name.log)"
  [ $rc -eq 0 ] && arms+=("$sw")
done
i=0
for arm in "${arms[@]}"; do
  i=$((i+1))
  envs=""; [ "$arm" != "-" ] && envs="$arm"
  ( env $envs timeout 90 python bench.py --dump-ops $out/ops_$i.json --no-cpu-baseline --no-eval-kernels --no-e2e ) \
      > $out/bench_$i.json 2> $out/bench_$i.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/bench_$i.json").read().strip().splitlines()[-1])
    print("== arm $i ($arm): ms_per_step %.3f  gemm %.3f ms  %s" % (d["ms_per_step"], d["roofline"]["per_kernel_ms"]["pdr_gemm_fused"],
          {k: round(v, 3) for k, v in d["roofline"]["per_kernel_ms"].items() if k != "pdr_gemm_fused"}))
except Exception as e:
    print("== arm $i ($arm): bench failed", e); print(open("$out/bench_$i.err").read()[-800:])
PY
done
