#!/bin/bash
# NMS A/B: gpurun --timeout 900 -- 'bash tools/gpu_nms.sh <tag>'
TAG=${1:-nms}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x --timeout 240 -k "nms" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log; tail -4 $OUT/pytest.log
for c0 in 512 256 128 64 32; do
YL_NMS_CHUNK0=$c0 timeout 300 python bench.py --no-extras --no-cpu-baseline --no-e2e --repeats 3 > $OUT/bench_c$c0.json 2> $OUT/bench_c$c0.err; python - <<PY
import json
d=json.loads(open("$OUT/bench_c$c0.json").read().strip().splitlines()[-1])
print("chunk0=$c0", {k:d[k] for k in ("value","value_serial","ms_per_step")}, d["bs1_latency_ms"]["p50"], "nms_select ms", d["kernel_breakdown"]["nms_select"]["ms"])
PY
done
