#!/bin/bash
# Round 2, GPU visit e: fused stage kernel after the epilogue rewrite -- its tests, a short bench, ncu (with source) of the
# first NFULL pdr_stage_chain launches of a step.
tag=${1:-r02e}
out=gpurun_out/$tag
mkdir -p $out
( timeout 600 python -m pytest tests/test_chain_gpu.py tests/test_model_gpu.py tests/test_ops_gpu.py tests/test_train_side.py -m gpu -q -s -x ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
grep -E "fused stages vs|passed|failed|FAILED|Error" $out/pytest_gpu.log | tail -12
( timeout 600 python bench.py --dump-ops $out/ops.json --no-gpu-reference --no-fast-ddpm --no-cpu-baseline ) > $out/bench.json 2> $out/bench.err; echo "bench exit $?"
tail -3 $out/bench.err
python - <<PY
import json
try:
    d = json.loads(open("$out/bench.json").read().strip().splitlines()[-1])
    print("== bench: ms_per_step %.3f  value %.3f e2e %.3f  roofline %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], {k: d["roofline"][k] for k in ("kernel", "frac", "per_kernel_ms")}))
    print("   emd", d["eval_kernels"]["emd_256x2048x2048"])
    ops = json.load(open("$out/ops.json"))
    tot = {}
    for o in ops:
        if o["op"] == "pdr_stage_chain":
            tot[o["stage"]] = tot.get(o["stage"], 0) + o["ms"]
            print("   chain %-10s sweep %d  %.4f ms  %.1f GFLOP  %.1f MB" % (o["stage"], o["sweep"], o["ms"], o["flops"] / 1e9, o["bytes"] / 1e6))
    print("   per stage", {k: round(v, 3) for k, v in tot.items()})
except Exception as e:
    print("bench parse failed", e)
PY
( timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:stage_chain \
    -c ${NFULL:-4} -f -o $out/chain_full \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-eval-kernels --no-gpu-reference --no-fast-ddpm --profiler-range ) > $out/ncu_full.log 2>&1
ncu -i $out/chain_full.ncu-rep --page raw --csv > $out/chain_full_raw.csv 2>/dev/null
ncu -i $out/chain_full.ncu-rep --page source --csv > $out/chain_full_source.csv 2>/dev/null
sz=$(stat -c %s $out/chain_full.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 45000000 ]; then rm -f $out/chain_full.ncu-rep; echo "report dropped ($sz bytes)" >> $out/ncu_full.log; fi
tail -2 $out/ncu_full.log | cut -c1-300
ls -la $out
