#!/bin/bash
# quick chain iteration: parity of the chain tests, timeline, layer table, short bench
TAG=${1:-chain}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -m pytest tests/test_fused_gpu.py -m gpu -q --timeout 120 -k "chain" > $OUT/pytest_chain.log 2>&1; echo "chain tests rc=$?"; tail -3 $OUT/pytest_chain.log
timeout 120 python tools/chain_timeline.py 64 > $OUT/chain_tl.txt 2>&1; head -24 $OUT/chain_tl.txt
timeout 120 python tools/layer_times.py n 64 > $OUT/layers_n64.txt 2>&1; grep -E "conv_chain|graph replay|launches" $OUT/layers_n64.txt | cut -c1-40,150-
timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-latency --no-extras > $OUT/bench.json 2> $OUT/bench.err; python - <<PY
import json
d=json.loads(open('$OUT/bench.json').read().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','value_serial')}, d.get('roofline',{}).get('frac'))
PY
