#!/bin/bash
# Quick GPU check: parity tests then the per-launch layer table.  gpurun --timeout 900 -- 'bash tools/gpu_quick.sh <tag> [pytest -k expr]'
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
KEXPR=${2:-}
if [ -n "$KEXPR" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q -k "$KEXPR" > $OUT/pytest.log 2>&1
else
  timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
fi
echo "pytest rc=$?" >> $OUT/pytest.log
tail -25 $OUT/pytest.log
timeout 300 python tools/layer_times.py n 64 > $OUT/layers_n64.txt 2>&1
tail -4 $OUT/layers_n64.txt
