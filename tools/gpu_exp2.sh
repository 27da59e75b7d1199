#!/bin/bash
TAG=${1:-e2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
./tools/mufu_bench.bin > $OUT/mufu.txt 2>&1; cat $OUT/mufu.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 240 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log; tail -8 $OUT/pytest.log
timeout 400 python bench.py --no-extras --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","value_serial","ms_per_step")}, d["bs1_latency_ms"]["p50"], d["e2e"]["value"], d["e2e_fp16_input"]["value"], d["e2e_u8_input"]["value"])
print(d["kernel_breakdown"])
PY
tail -3 $OUT/bench.err
