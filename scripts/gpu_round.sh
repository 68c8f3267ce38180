#!/bin/bash
# One consolidated GPU-box visit: parity tests, smoke, both bench arms, ncu launch list, ncu full capture of the GEMMs.
# Usage (from the repo root, through gpurun): bash scripts/gpu_round.sh <tag>
tag=${1:-rXX}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke ) > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log
( timeout 600 python bench.py --dump-ops $out/ops.json ) > $out/bench.json 2> $out/bench.err
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ) > $out/bench_reference.json 2> $out/bench_reference.err
if [ -z "$SKIP_NCU" ]; then
( timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-eval-kernels --profiler-range ) > $out/ncu_launches.log 2>&1
( timeout 480 ncu --profile-from-start off --set full --clock-control none -k regex:gemm -f -o $out/gemm_full \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-eval-kernels --profiler-range ) > $out/ncu_full.log 2>&1
fi
tail -3 $out/pytest_gpu.log; tail -2 $out/smoke.log; cat $out/bench.json; cat $out/bench_reference.json
ls -la $out
