#!/bin/bash
# ncu stall sampling of the gathered-A GEMMs (scripts/prof_gather.py), cp.async pieces vs TMA gather4
out=gpurun_out/${1:-r02s}
mkdir -p $out
for arm in 0 1; do
  PDR_GEMM_TMA_GATHER=$arm timeout 120 python scripts/prof_gather.py
  PDR_GEMM_TMA_GATHER=$arm timeout 600 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section SpeedOfLight \
    --section MemoryWorkloadAnalysis --import-source on --clock-control none -k regex:gemm_tf32 -c 6 -o $out/gather_$arm -f \
    python scripts/prof_gather.py > $out/ncu_$arm.log 2>&1
  tail -3 $out/ncu_$arm.log
done
ls -la $out
