#!/bin/bash
# layer tables + quick bench: gpurun --timeout 900 -- 'bash tools/gpu_layers.sh <tag> [pytest -k expr]'
TAG=${1:-l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -n "$2" ]; then timeout 600 python -m pytest tests -m gpu -q -x --timeout 240 -k "$2" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log; tail -6 $OUT/pytest.log; fi
timeout 120 python tools/layer_times.py n 64 > $OUT/layers_n64.txt 2>&1; tail -1 $OUT/layers_n64.txt
timeout 120 python tools/layer_times.py n 1 > $OUT/layers_n1.txt 2>&1; tail -1 $OUT/layers_n1.txt
timeout 180 python tools/layer_times.py m 64 > $OUT/layers_m64.txt 2>&1; tail -1 $OUT/layers_m64.txt
timeout 400 python bench.py --no-extras --no-cpu-baseline --no-e2e > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","value_serial","ms_per_step")}, d["bs1_latency_ms"]["p50"])
print({k:v["ms"] for k,v in d["kernel_breakdown"].items()})
PY
tail -3 $OUT/bench.err
