#!/bin/bash
# Short end-of-round visit (GPU budget nearly spent): parity tests, one bench line (CPU arm skipped), ncu launch list.
tag=${1:-rXXshort}
out=gpurun_out/$tag
mkdir -p $out
( timeout 150 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -5 $out/pytest_gpu.log
( timeout 60 python __graft_entry__.py smoke ) > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log; tail -2 $out/smoke.log
( timeout 120 python bench.py --no-cpu-baseline --dump-ops $out/ops.json ) > $out/bench.json 2> $out/bench.err
tail -c 2500 $out/bench.json
( timeout 90 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-eval-kernels --profiler-range ) > $out/ncu_launches.log 2>&1
wc -l $out/launches.csv
