#!/bin/bash
# One gpurun call: parity tests, layer table, bench, ncu launch list + full capture of the conv kernel.
# Usage (here): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
timeout 300 python tools/layer_times.py n 64 > $OUT/layers_n64.txt 2>&1
timeout 300 python tools/layer_times.py m 64 > $OUT/layers_m64.txt 2>&1
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -c 3000 $OUT/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 80 -c 24 -f -o $OUT/conv_tc \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
ls -la $OUT
