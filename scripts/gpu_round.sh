#!/bin/bash
# One consolidated GPU-box visit: parity tests, smoke, both bench arms, ncu launch list, ncu full capture of GEMMs.
# Usage (from the repo root, through gpurun): bash scripts/gpu_round.sh <tag> [pytest args]
# gpurun copies back at most 64 MiB of gpurun_out/: reports are exported to CSV on the box and large ones dropped.
tag=${1:-rXX}
shift
pytest_args=${@:-tests -m gpu}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
( timeout 900 python -m pytest $pytest_args -x -q ) > $out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> $out/pytest_gpu.log
COLD=""
if [ $rc -ne 0 ]; then COLD="--module-cold"; fi      # compiled cold path only when its parity tests are green
( timeout 300 python __graft_entry__.py smoke ) > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log
( timeout 900 python bench.py $COLD --dump-ops $out/ops.json ) > $out/bench.json 2> $out/bench.err
if [ -z "$SKIP_REF" ]; then
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ) > $out/bench_reference.json 2> $out/bench_reference.err
fi
if [ -z "$SKIP_NCU" ]; then
( timeout 420 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $out/launches.csv \
    python bench.py $COLD --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-eval-kernels --profiler-range ) > $out/ncu_launches.log 2>&1
# full capture (with source) of the first NFULL GEMM launches of a step = the encoder mapper block at 2 M rows:
# every producer mode (direct, GN/ReLU transform, residual, broadcast row-add)
( timeout 420 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:gemm_tf32 \
    -c ${NFULL:-12} -f -o $out/gemm_full \
    python bench.py $COLD --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-eval-kernels --profiler-range ) > $out/ncu_full.log 2>&1
ncu -i $out/gemm_full.ncu-rep --page raw --csv > $out/gemm_full_raw.csv 2>/dev/null
( timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -f -o $out/ops_full \
    python scripts/ncu_ops.py ) > $out/ncu_ops.log 2>&1
ncu -i $out/ops_full.ncu-rep --page raw --csv > $out/ops_full_raw.csv 2>/dev/null
osz=$(stat -c %s $out/ops_full.ncu-rep 2>/dev/null || echo 0)
if [ "$osz" -gt 15000000 ]; then rm -f $out/ops_full.ncu-rep; fi
sz=$(stat -c %s $out/gemm_full.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 30000000 ]; then rm -f $out/gemm_full.ncu-rep; echo "report dropped ($sz bytes)" >> $out/ncu_full.log; fi
fi
tail -3 $out/pytest_gpu.log; tail -2 $out/smoke.log; cat $out/bench.json
[ -z "$SKIP_REF" ] && cat $out/bench_reference.json
tail -3 $out/ncu_launches.log $out/ncu_full.log 2>/dev/null
du -sh gpurun_out; ls -la $out
