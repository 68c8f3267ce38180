#!/bin/bash
# ncu stall sampling of the real engine's GEMMs: ops 258..266 (last grouped stage, 2 M rows) and 239..247 (524288 rows)
out=gpurun_out/${1:-r02v}
mkdir -p $out
for rng in "258 267" "239 248"; do
  set -- $rng
  timeout 900 ncu --profile-from-start off --section SourceCounters --section WarpStateStats --section SchedulerStats --section SpeedOfLight \
    --section MemoryWorkloadAnalysis --import-source on --clock-control none -k regex:gemm_tf32 -o $out/ops_$1 -f \
    python scripts/prof_engine_ops.py $1 $2 > $out/ncu_$1.log 2>&1
  grep -E "^ops|done|Error|error" $out/ncu_$1.log | cut -c1-400
done
ls -la $out
