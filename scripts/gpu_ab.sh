#!/bin/bash
# A/B of the GEMM epilogue flavours on one box: parity tests, then bench with each (short form for the B arm).
tag=${1:-ab}
out=gpurun_out/$tag
mkdir -p $out
( timeout 900 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
( PDR_GEMM_EPILOGUE=scalar timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py -m gpu -x -q ) > $out/pytest_scalar.log 2>&1; echo "pytest exit $?" >> $out/pytest_scalar.log
( timeout 600 python bench.py --dump-ops $out/ops_vec.json --no-cpu-baseline ) > $out/bench_vec.json 2> $out/bench_vec.err
( PDR_GEMM_EPILOGUE=scalar timeout 600 python bench.py --dump-ops $out/ops_scalar.json --no-cpu-baseline --no-eval-kernels --no-e2e ) > $out/bench_scalar.json 2> $out/bench_scalar.err
tail -4 $out/pytest_gpu.log; tail -2 $out/pytest_scalar.log; cat $out/bench_vec.json; cat $out/bench_scalar.json; tail -5 $out/*.err
