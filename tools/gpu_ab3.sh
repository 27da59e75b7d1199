#!/bin/bash
# A/B of schedule-level knobs on the default bench (value only)
OUT=gpurun_out/${1:-ab3}
mkdir -p $OUT
run() {
  tag=$1; shift
  env "$@" timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-latency --no-extras --no-roofline $EXTRA > $OUT/$tag.json 2> $OUT/$tag.err
  python - <<PY
import json
try:
    d=json.loads(open('$OUT/$tag.json').read().splitlines()[-1])
    print('$tag', {k:d.get(k) for k in ('value','ms_per_step','value_serial')})
except Exception as e:
    print('$tag', 'failed', e)
PY
}
EXTRA="--inflight 2" run base X=1
EXTRA="--inflight 3" run fly3 X=1
EXTRA="--inflight 4" run fly4 X=1
EXTRA="--inflight 2" run grid1 YL_GRID_CTAS=1
EXTRA="--inflight 3" run grid1_fly3 YL_GRID_CTAS=1
EXTRA="--inflight 2" run nopdl YL_PDL=0
EXTRA="--inflight 2" run nolanes YL_DET_LANES=0
