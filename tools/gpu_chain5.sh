#!/bin/bash
# chain on / off at small batches with batch-dependent cluster sizes
OUT=gpurun_out/${1:-chaincs}
mkdir -p $OUT
timeout 300 python -m pytest tests/test_fused_gpu.py -m gpu -q --timeout 120 -k "chain" 2>&1 | tail -3
for b in 1 2 4 8 16 32; do
  r0=$(YL_CHAIN=0 timeout 120 python tools/layer_times.py n $b 2>&1 | tail -1)
  echo "bs=$b chain=0: $r0"
  for cs in 0 8 4; do
    r=$(YL_CHAIN=1 YL_CHAIN_CLUSTER=$cs timeout 120 python tools/layer_times.py n $b 2>&1 | tee $OUT/layers_n${b}_cs$cs.txt | tail -1)
    echo "bs=$b chain=1 cluster=$cs: $r"
  done
done
