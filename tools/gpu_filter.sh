#!/bin/bash
OUT=gpurun_out/${1:-filt}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_nms_gpu.py tests/test_real_images.py -m gpu -q --timeout 200 > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest.log
timeout 120 python tools/layer_times.py n 64 > $OUT/layers_n64.txt 2>&1; grep -E "filter|graph replay|launches" $OUT/layers_n64.txt
timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-latency --no-extras > $OUT/bench.json 2> $OUT/bench.err; python - <<PY
import json
d=json.loads(open('$OUT/bench.json').read().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','value_serial')}, d.get('roofline',{}).get('frac'))
PY
