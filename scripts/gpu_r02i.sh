#!/bin/bash
# A/B of the fused stage kernel's epilogue width: 4 vs 8 warps per tile group (PDR_CHAIN_WPG), chain tests first.
tag=${1:-r02i}
out=gpurun_out/$tag
mkdir -p $out
( timeout 600 python -m pytest tests/test_chain_gpu.py tests/test_dropin_gpu.py tests/test_train_side.py -m gpu -q -x ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -3 $out/pytest_gpu.log
for w in 4 8; do
  ( PDR_CHAIN_WPG=$w timeout 300 python bench.py --dump-ops $out/ops_$w.json --no-gpu-reference --no-fast-ddpm --no-cpu-baseline --no-eval-kernels --no-e2e ) > $out/bench_$w.json 2> $out/bench_$w.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/bench_$w.json").read().strip().splitlines()[-1])
    ops = json.load(open("$out/ops_$w.json"))
    tot = {}
    for o in ops:
        if o["op"] == "pdr_stage_chain":
            tot[o["stage"]] = tot.get(o["stage"], 0) + o["ms"]
    print("== WPG=$w: ms_per_step %.3f chain %.3f ms per stage %s" % (d["ms_per_step"], d["roofline"]["per_kernel_ms"].get("pdr_stage_chain", 0), {k: round(v, 3) for k, v in tot.items()}))
    print("   sweeps enc_map0", [round(o["ms"], 3) for o in ops if o["op"] == "pdr_stage_chain" and o["stage"] == "enc_map0"])
except Exception as e:
    print("bench parse failed", e); print(open("$out/bench_$w.err").read()[-600:])
PY
done
