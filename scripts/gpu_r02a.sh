#!/bin/bash
# Round 2, first GPU visit: full parity suite (with the new B=32 / TF32-distribution / FastDPM-loop tests), smoke, the
# extended bench line, then the A/B of the four round-1 experiments (scripts/gpu_round2_ab.sh).
tag=${1:-r02a}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -q -s -x ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
grep -E "TF32|cd_t\(|passed|failed|Error|error" $out/pytest_gpu.log | tail -30
( timeout 300 python __graft_entry__.py smoke ) > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log
tail -2 $out/smoke.log
( timeout 600 python bench.py --dump-ops $out/ops.json ) > $out/bench.json 2> $out/bench.err; echo "bench exit $?"
tail -5 $out/bench.err
cat $out/bench.json
bash scripts/gpu_round2_ab.sh $tag/ab
