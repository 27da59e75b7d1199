#!/bin/bash
# only the ncu launch-list pass of the profile round (per-launch time + DRAM bytes of every kernel of the bench's steps)
OUT=gpurun_out/${1:-ll}
mkdir -p $OUT
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --inflight 1 --repeats 1 --no-cpu-baseline --no-e2e --no-latency --no-roofline --no-extras > $OUT/ncu_bench.log 2>&1
grep -c conv_tc $OUT/launches.csv; tail -2 $OUT/ncu_bench.log | cut -c1-300
