#!/bin/bash
# Round 2, GPU visit c: first contact of the fused stage kernel (pdr_stage_chain).
#   1. its unit tests against the numpy emulator, under a hard timeout (a deadlocked kernel must not eat the box)
#   2. the full parity suite + bench with the fused stages on (only if 1 passed), else with PDR_STAGE_CHAIN=0
#   3. the four round-1 experiments on the per-layer engine (PDR_STAGE_CHAIN=0), each: parity tests, then a bench arm
tag=${1:-r02c}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
( timeout 300 python -m pytest tests/test_chain_gpu.py -m gpu -q -s -x ) > $out/pytest_chain.log 2>&1; rc=$?
echo "chain pytest exit $rc" | tee -a $out/pytest_chain.log
grep -E "fused stages vs|passed|failed|Error|error|assert" $out/pytest_chain.log | tail -25
if [ $rc -ne 0 ]; then export PDR_STAGE_CHAIN=0; echo "== falling back to PDR_STAGE_CHAIN=0 for the rest"; fi
( timeout 900 python -m pytest tests -m gpu -q -s --deselect tests/test_chain_gpu.py ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
grep -E "TF32|cd_t\(|passed|failed|FAILED" $out/pytest_gpu.log | tail -30
( timeout 300 python __graft_entry__.py smoke ) > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log
tail -2 $out/smoke.log
( timeout 600 python bench.py --dump-ops $out/ops.json --no-gpu-reference --no-fast-ddpm --no-cpu-baseline --no-eval-kernels ) > $out/bench.json 2> $out/bench.err; echo "bench exit $?"
tail -3 $out/bench.err
python - <<PY
import json
try:
    d = json.loads(open("$out/bench.json").read().strip().splitlines()[-1])
    print("== bench: ms_per_step %.3f  roofline %s" % (d["ms_per_step"], {k: d["roofline"][k] for k in ("kernel", "frac", "per_kernel_ms")}))
    ops = json.load(open("$out/ops.json"))
    for o in ops:
        if o["op"] == "pdr_stage_chain":
            print("   chain %-10s sweep %d  %.4f ms  %.1f GFLOP  %.1f MB" % (o["stage"], o["sweep"], o["ms"], o["flops"] / 1e9, o["bytes"] / 1e6))
except Exception as e:
    print("bench parse failed", e)
PY
export PDR_STAGE_CHAIN=0
bash scripts/gpu_round2_ab.sh $tag/ab
