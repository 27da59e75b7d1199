#!/bin/bash
# A/B round: parity tests, then the layer table under env-switch variants, then an ncu source capture of a
# named kernel.  gpurun --timeout 1200 -- 'bash tools/gpu_ab.sh <tag> "<VAR=val VAR=val;VAR=val>" [ncu kernel regex]'
TAG=${1:-ab}
VARIANTS=${2:-}
KREGEX=${3:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -x -q --timeout 90 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -15 $OUT/pytest.log
timeout 120 python tools/layer_times.py n 64 > $OUT/layers_n64.txt 2>&1; tail -1 $OUT/layers_n64.txt
IFS=';' read -ra VS <<< "$VARIANTS"
i=0
for v in "${VS[@]}"; do
  i=$((i+1))
  echo "variant $i: $v"
  env $v timeout 120 python tools/layer_times.py n 64 > $OUT/layers_n64_v$i.txt 2>&1; echo "# $v" >> $OUT/layers_n64_v$i.txt; tail -2 $OUT/layers_n64_v$i.txt
done
timeout 120 python tools/bench_kernels.py dwconv > $OUT/bench_kernels.txt 2>&1; cat $OUT/bench_kernels.txt
timeout 240 python bench.py --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; tail -c 1500 $OUT/bench.json
if [ -n "$KREGEX" ]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 3 -c 2 -f -o $OUT/kern \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_kern.log 2>&1
fi
ls -la $OUT
