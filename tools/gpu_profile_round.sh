#!/bin/bash
# Profile round: bench (ours n64 / m256 / reference arm), layer tables, ncu launch list + DRAM traffic, ncu --set full of
# the main kernels.   gpurun --timeout 1500 -- 'bash tools/gpu_profile_round.sh <tag>'
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 500 python -m pytest tests -m gpu -x -q --timeout 120 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log; tail -2 $OUT/pytest.log
timeout 300 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -c 600 $OUT/bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
timeout 300 python bench.py --model m --batch 256 --steps 10 --no-cpu-baseline --no-latency > $OUT/bench_m256.json 2> $OUT/bench_m256.err
timeout 120 python tools/layer_times.py n 64 > $OUT/layers_n64.txt 2>&1; tail -1 $OUT/layers_n64.txt
timeout 180 python tools/layer_times.py m 64 > $OUT/layers_m64.txt 2>&1; tail -1 $OUT/layers_m64.txt
timeout 120 python tools/layer_times.py n 1 > $OUT/layers_n1.txt 2>&1; tail -1 $OUT/layers_n1.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --inflight 1 --repeats 1 --no-cpu-baseline --no-e2e --no-latency --no-roofline --no-extras > $OUT/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel" -s 240 -c 12 -f -o $OUT/conv_tc \
    python bench.py --steps 1 --warmup 3 --inflight 1 --no-cpu-baseline --no-e2e --no-latency > $OUT/ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"c3k2_tail|dwconv3x3_mma|dwpw_tc|stem_|letterbox|nms_|psa_att|sppf" -s 12 -c 16 -f -o $OUT/others \
    python bench.py --steps 1 --warmup 3 --inflight 1 --no-cpu-baseline --no-latency > $OUT/ncu_others.log 2>&1
ls -la $OUT
