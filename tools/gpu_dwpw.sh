#!/bin/bash
OUT=gpurun_out/${1:-dwpw}
mkdir -p $OUT
timeout 300 python -m pytest tests/test_fused_gpu.py -m gpu -q --timeout 120 -x -k "dw_pw or fused_dw" > $OUT/pytest_dwpw.log 2>&1; echo "dwpw tests rc=$?"; tail -12 $OUT/pytest_dwpw.log
timeout 120 python tools/layer_times.py n 64 > $OUT/layers_n64.txt 2>&1; grep -E "dwpw|dwconv|graph replay|launches" $OUT/layers_n64.txt
timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-latency --no-extras > $OUT/bench.json 2> $OUT/bench.err; python - <<PY
import json
d=json.loads(open('$OUT/bench.json').read().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','value_serial')}, d.get('roofline',{}).get('frac'))
PY
YL_DWPW=0 timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-latency --no-extras > $OUT/bench_off.json 2> $OUT/bench_off.err; python - <<PY
import json
d=json.loads(open('$OUT/bench_off.json').read().splitlines()[-1])
print('off', {k:d.get(k) for k in ('value','ms_per_step','value_serial')}, d.get('roofline',{}).get('frac'))
PY
