#!/bin/bash
out=gpurun_out/${1:-chamfer}
mkdir -p $out
( timeout 600 python -m pytest tests -m gpu -q -k "chamfer or Chamfer or drop" ) > $out/pytest_chamfer.log 2>&1; tail -3 $out/pytest_chamfer.log; grep -E "^E " $out/pytest_chamfer.log | head
timeout 200 python scripts/chamfer_probe.py 2>&1 | tail -1 | tee $out/chamfer.json
