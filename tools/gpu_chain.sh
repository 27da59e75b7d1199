#!/bin/bash
# Conv-chain check: parity tests of the chain kernel, layer table and a short bench with / without chains.
# gpurun --timeout 900 -- 'bash tools/gpu_chain.sh <tag>'
TAG=${1:-chain}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -m pytest tests/test_fused_gpu.py -m gpu -q --timeout 120 -k "chain" > $OUT/pytest_chain.log 2>&1; echo "chain tests rc=$?"; tail -15 $OUT/pytest_chain.log
timeout 120 python tools/layer_times.py n 64 > $OUT/layers_n64.txt 2>&1; tail -1 $OUT/layers_n64.txt
YL_CHAIN=0 timeout 120 python tools/layer_times.py n 64 > $OUT/layers_n64_nochain.txt 2>&1; tail -1 $OUT/layers_n64_nochain.txt
timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-latency --no-extras > $OUT/bench.json 2> $OUT/bench.err; python - <<PY
import json
d=json.loads(open('$OUT/bench.json').read().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','value_serial')}, d.get('roofline',{}).get('frac'))
PY
YL_CHAIN=0 timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-latency --no-extras > $OUT/bench_nochain.json 2> $OUT/bench_nochain.err; python - <<PY
import json
d=json.loads(open('$OUT/bench_nochain.json').read().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','value_serial')}, d.get('roofline',{}).get('frac'))
PY
