#!/bin/bash
OUT=gpurun_out/${1:-b2bbox}
mkdir -p $OUT
timeout 300 python -m pytest tests/test_fused_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 120 -x -k "back_to_back or fused_decode or infer_nms or model_640" > $OUT/pytest.log 2>&1; echo "tests rc=$?"; tail -12 $OUT/pytest.log
timeout 120 python tools/layer_times.py n 64 > $OUT/layers_n64.txt 2>&1; grep -E "decode|graph replay|launches" $OUT/layers_n64.txt | cut -c1-130
for m in 1 0; do
YL_CONV_DET=$m timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-latency --no-extras > $OUT/bench_$m.json 2> $OUT/bench_$m.err; python - <<PY
import json
d=json.loads(open('$OUT/bench_$m.json').read().splitlines()[-1])
print('conv_det=$m', {k:d.get(k) for k in ('value','ms_per_step','value_serial','launches_per_step','detections_last_step')}, d.get('roofline',{}).get('frac'))
PY
done
