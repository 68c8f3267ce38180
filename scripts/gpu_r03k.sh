#!/bin/bash
# final build: ncu --set full of the last grouped stage's GEMMs (ops 258..266) + csv summary; 2-GPU run of both arms
out=gpurun_out/${1:-r03k}
mkdir -p $out
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:gemm_tf32 -o $out/gemm_last_stage_full -f \
    python scripts/prof_engine_ops.py 258 267 > $out/ncu_full.log 2>&1
tail -2 $out/ncu_full.log
ncu -i $out/gemm_last_stage_full.ncu-rep --page raw --csv > $out/gemm_last_stage_full_raw.csv 2>/dev/null
ls -la $out
