#!/bin/bash
# Round 2, GPU visit d: fused stages on by default.  Parity suite, smoke, full bench line, ncu launch list of a step,
# ncu full capture (with source) of the pdr_stage_chain launches of one step.
tag=${1:-r02d}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -q -s ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
grep -E "TF32|cd_t\(|fused stages vs|passed|failed|FAILED" $out/pytest_gpu.log | tail -30
( timeout 300 python __graft_entry__.py smoke ) > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log
tail -2 $out/smoke.log
( timeout 900 python bench.py --dump-ops $out/ops.json ) > $out/bench.json 2> $out/bench.err; echo "bench exit $?"
tail -3 $out/bench.err
python - <<PY
import json
try:
    d = json.loads(open("$out/bench.json").read().strip().splitlines()[-1])
    print("== bench: ms_per_step %.3f  value %.3f e2e %.3f  roofline %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], {k: d["roofline"][k] for k in ("kernel", "frac", "per_kernel_ms")}))
    print("   fp32_simt", d.get("fp32_simt"), "gpu_reference", d.get("gpu_reference"))
    print("   fast_ddpm", d.get("fast_ddpm"))
    print("   eval", d.get("eval_kernels"))
    ops = json.load(open("$out/ops.json"))
    for o in ops:
        if o["op"] == "pdr_stage_chain":
            print("   chain %-10s sweep %d  %.4f ms  %.1f GFLOP  %.1f MB" % (o["stage"], o["sweep"], o["ms"], o["flops"] / 1e9, o["bytes"] / 1e6))
except Exception as e:
    print("bench parse failed", e)
PY
( timeout 420 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-eval-kernels --no-gpu-reference --no-fast-ddpm --profiler-range ) > $out/ncu_launches.log 2>&1
( timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:stage_chain \
    -c 21 -f -o $out/chain_full \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-eval-kernels --no-gpu-reference --no-fast-ddpm --profiler-range ) > $out/ncu_full.log 2>&1
ncu -i $out/chain_full.ncu-rep --page raw --csv > $out/chain_full_raw.csv 2>/dev/null
sz=$(stat -c %s $out/chain_full.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 40000000 ]; then rm -f $out/chain_full.ncu-rep; echo "report dropped ($sz bytes)" >> $out/ncu_full.log; fi
tail -3 $out/ncu_launches.log $out/ncu_full.log 2>/dev/null
du -sh gpurun_out/$tag; ls -la $out
