#!/bin/bash
# chain on/off across batch sizes (graph replay of the model) + the bs=64 bench
OUT=gpurun_out/${1:-chainsweep}
mkdir -p $OUT
for b in 4 8 16 32 64; do
  for c in 0 1; do
    r=$(YL_CHAIN=$c YL_CHAIN_MIN_BATCH=1 timeout 120 python tools/layer_times.py n $b 2>&1 | tee $OUT/layers_n${b}_c$c.txt | tail -1)
    echo "bs=$b chain=$c: $r"
  done
done
timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-latency --no-extras > $OUT/bench.json 2> $OUT/bench.err; python - <<PY
import json
d=json.loads(open('$OUT/bench.json').read().splitlines()[-1])
print('chain on', {k:d.get(k) for k in ('value','ms_per_step','value_serial')}, d.get('roofline',{}).get('frac'))
PY
