#!/bin/bash
# bench only (+ optional layer tables): gpurun --timeout 900 -- 'bash tools/gpu_bench.sh <tag> [layers]'
TAG=${1:-b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 7000 $OUT/bench.json; tail -5 $OUT/bench.err
if [ -n "$2" ]; then
  timeout 120 python tools/layer_times.py n 64 > $OUT/layers_n64.txt 2>&1; tail -1 $OUT/layers_n64.txt
  timeout 120 python tools/layer_times.py n 1 > $OUT/layers_n1.txt 2>&1; tail -1 $OUT/layers_n1.txt
  timeout 120 python tools/timeline.py n 1 > $OUT/timeline_n1.txt 2>&1; tail -3 $OUT/timeline_n1.txt
fi
