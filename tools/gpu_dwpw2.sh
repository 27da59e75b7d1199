#!/bin/bash
OUT=gpurun_out/${1:-dwpwk}
mkdir -p $OUT
for cfg in "2 2" "3 2" "4 2" "3 3" "4 3"; do
  set -- $cfg
  echo "== raw=$1 a=$2"
  YL_DWPW_RAW=$1 YL_DWPW_A=$2 timeout 120 python tools/layer_times.py n 64 2>&1 | grep -E "dwpw_tc|graph replay" | awk '{print $1,$2,$(NF-3)}' | tr '\n' ';'; echo
done
