#!/bin/bash
# A/B of the chain knobs: timeline of the first chain under patch / resident-weight settings
OUT=gpurun_out/${1:-chainab}
mkdir -p $OUT
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  echo "=== YL_CHAIN_PATCH=$1 YL_CHAIN_BRES=$2"
  YL_CHAIN_PATCH=$1 YL_CHAIN_BRES=$2 timeout 120 python tools/chain_timeline.py 64 > $OUT/tl_p$1_b$2.txt 2>&1
  head -21 $OUT/tl_p$1_b$2.txt | cut -c1-100
done
timeout 300 python -m pytest tests/test_fused_gpu.py -m gpu -q --timeout 120 -k "chain" 2>&1 | tail -3
