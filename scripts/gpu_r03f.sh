#!/bin/bash
# Round 2, final default engine (PDL, balanced + narrow column tiles, lean gathered producer, packed prologue / statistics).
# Full parity suite, smoke, full bench line, EMD cluster-size A/B, ncu launch list of one step.
tag=${1:-r03f}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -q -s ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
grep -E "TF32|cd_t\(|passed|failed|FAILED" $out/pytest_gpu.log | tail -14
( timeout 300 python __graft_entry__.py smoke ) > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log
tail -2 $out/smoke.log
( timeout 900 python bench.py --dump-ops $out/ops.json ) > $out/bench.json 2> $out/bench.err; echo "bench exit $?"
tail -3 $out/bench.err
python - <<PY
import json
try:
    d = json.loads(open("$out/bench.json").read().strip().splitlines()[-1])
    print("== bench: ms_per_step %.3f  value %.3f e2e %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["value"]))
    print("   roofline", {k: d["roofline"][k] for k in ("kernel", "frac", "tensor_frac", "tensor_core_kernels", "per_kernel_ms")})
    print("   fp32_simt", d.get("fp32_simt"), "gpu_reference", {k: v for k, v in (d.get("gpu_reference") or {}).items() if k != "what"})
    print("   fast_ddpm", d.get("fast_ddpm"))
    print("   eval", d.get("eval_kernels"))
except Exception as e:
    print("bench parse failed", e)
PY
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $out/bench_reference.json 2> $out/bench_reference.err; tail -c 400 $out/bench_reference.json
( timeout 420 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-eval-kernels --no-gpu-reference --no-fast-ddpm --profiler-range ) > $out/ncu_launches.log 2>&1
tail -2 $out/ncu_launches.log | cut -c1-200
ls -la $out
