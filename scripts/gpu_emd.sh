#!/bin/bash
# EMD kernel: parity tests, then timing at B = 256 / 32 (auto cluster) with checksums
out=gpurun_out/${1:-emd}
mkdir -p $out
( timeout 600 python -m pytest tests -m gpu -q -k "emd or EMD" ) > $out/pytest_emd.log 2>&1; tail -3 $out/pytest_emd.log; grep -E "^E " $out/pytest_emd.log | head
timeout 120 python scripts/emd_cluster_ab.py 2>&1 | tail -1 | tee $out/emd_auto.json
PDR_EMD_CLUSTER=4 timeout 120 python scripts/emd_cluster_ab.py 2>&1 | tail -1 | tee $out/emd_c4.json
