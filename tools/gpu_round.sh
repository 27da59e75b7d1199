#!/bin/bash
# One gpurun call: parity tests, layer tables, bench, ncu launch list + DRAM traffic + full capture of the conv kernel.
# Usage (here): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 400 python -m pytest tests -m gpu -x -q --timeout 90 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
timeout 120 python tools/layer_times.py n 64 > $OUT/layers_n64.txt 2>&1; tail -1 $OUT/layers_n64.txt
timeout 180 python tools/layer_times.py m 64 > $OUT/layers_m64.txt 2>&1; tail -1 $OUT/layers_m64.txt
timeout 120 python tools/layer_times.py n 1 > $OUT/layers_n1.txt 2>&1; tail -1 $OUT/layers_n1.txt
timeout 300 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -c 2500 $OUT/bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; tail -c 600 $OUT/bench_ref.json
timeout 300 python bench.py --model m --batch 256 --steps 10 --no-cpu-baseline > $OUT/bench_m256.json 2> $OUT/bench_m256.err; tail -c 1200 $OUT/bench_m256.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 240 -c 12 -f -o $OUT/conv_tc \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
ls -la $OUT
