#!/bin/bash
OUT=gpurun_out/${1:-ab4}
mkdir -p $OUT
run() {
  tag=$1; shift
  env "$@" timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-latency --no-extras --no-roofline > $OUT/$tag.json 2> $OUT/$tag.err
  python - <<PY
import json
try:
    d=json.loads(open('$OUT/$tag.json').read().splitlines()[-1])
    print('$tag', {k:d.get(k) for k in ('value','ms_per_step','value_serial')})
except Exception as e:
    print('$tag', 'failed', e)
PY
}
run base X=1
run nolanes YL_DET_LANES=0
run convdet YL_CONV_DET=1
run convdet_nolanes YL_CONV_DET=1 YL_DET_LANES=0
