#!/bin/bash
# Quick A/B on one box: parity tests, then short bench arms selected by environment switches.
# usage: [PYTEST_ENV="ENV=.."] [ARM_TIMEOUT=s] bash scripts/gpu_ab.sh <tag> "<pytest args>" "ENV1=.. ENV2=.." "ENVb=.." ...
#        (one bench arm per extra argument; "-" = default env; PYTEST_ENV is applied to the pytest run only)
tag=${1:-ab}; shift
pyt=${1:-tests -m gpu}; shift
out=gpurun_out/$tag
mkdir -p $out
( env $PYTEST_ENV timeout ${PYTEST_TIMEOUT:-900} python -m pytest $pyt -x -q ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -${PYTEST_TAIL:-4} $out/pytest_gpu.log
i=0
for arm in "$@"; do
  i=$((i+1))
  envs=""; [ "$arm" != "-" ] && envs="$arm"
  ( env $envs timeout ${ARM_TIMEOUT:-600} python bench.py --dump-ops $out/ops_$i.json --no-cpu-baseline --no-eval-kernels --no-e2e $BENCH_ARGS ) > $out/bench_$i.json 2> $out/bench_$i.err
  echo "== arm $i: $arm"; python - <<PY
import json
try:
    d=json.loads(open("$out/bench_$i.json").read().strip().splitlines()[-1])
    print("ms_per_step %.3f  frac %.3f  %s" % (d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["per_kernel_ms"]))
except Exception as e:
    print("bench failed", e); print(open("$out/bench_$i.err").read()[-1500:])
PY
done
if [ -n "$BW_PROBE" ]; then python scripts/bw_probe.py | tee $out/bw_probe.json; fi
