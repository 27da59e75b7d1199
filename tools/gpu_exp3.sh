#!/bin/bash
TAG=${1:-e3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 240 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log; tail -12 $OUT/pytest.log
timeout 120 python tools/layer_times.py n 64 > $OUT/layers_n64.txt 2>&1; tail -1 $OUT/layers_n64.txt
YL_PATCH_EFF=50 timeout 120 python tools/layer_times.py n 64 > $OUT/layers_n64_eff50.txt 2>&1; tail -1 $OUT/layers_n64_eff50.txt
YL_PATCH_EFF=50 timeout 120 python tools/layer_times.py n 1 > $OUT/layers_n1_eff50.txt 2>&1; tail -1 $OUT/layers_n1_eff50.txt
timeout 120 python tools/layer_times.py n 1 > $OUT/layers_n1.txt 2>&1; tail -1 $OUT/layers_n1.txt
for cw in 1 0; do
YL_NMS_CLASSWISE=$cw timeout 400 python bench.py --no-extras --no-cpu-baseline --no-e2e > $OUT/bench_cw$cw.json 2> $OUT/bench_cw$cw.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("$OUT/bench_cw$cw.json").read().strip().splitlines()[-1])
print("classwise=$cw", {k:d[k] for k in ("value","value_serial","ms_per_step")}, d["bs1_latency_ms"]["p50"])
print({k:v["ms"] for k,v in d["kernel_breakdown"].items()})
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stem_conv -c 1 -f -o $OUT/stem_conv python tools/layer_times.py m 16 > $OUT/ncu_stem.log 2>&1; tail -2 $OUT/ncu_stem.log
ls -la $OUT | head -20
