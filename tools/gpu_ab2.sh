#!/bin/bash
TAG=${1:-ab2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
YL_C3K2_FUSE=0 timeout 120 python tools/layer_times.py n 64 > $OUT/layers_n64_nofusetail.txt 2>&1; tail -1 $OUT/layers_n64_nofusetail.txt
YL_STEM_FUSE=0 timeout 120 python tools/layer_times.py n 64 > $OUT/layers_n64_nofusestem.txt 2>&1; tail -1 $OUT/layers_n64_nofusestem.txt
timeout 120 python tools/layer_times.py n 64 > $OUT/layers_n64.txt 2>&1; tail -1 $OUT/layers_n64.txt
head -12 $OUT/layers_n64_nofusetail.txt
