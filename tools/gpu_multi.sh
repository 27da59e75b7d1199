#!/bin/bash
# multi-GPU bench exactly as the driver launches it: gpurun --gpus N --timeout 900 -- 'bash tools/gpu_multi.sh <tag> N'
TAG=${1:-mg}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "bench rc=$?"; tail -c 400 $OUT/bench_n$N.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_n$N.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","n_gpus","ms_per_step")})
for k in ("e2e","e2e_fp16_input","e2e_u8_input"): print(k, d[k]["value"])
print(d["h2d_ceiling"]); print(d["yolo11m_bs256"]); print(d["clocks"])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $OUT/ref_n$N.json 2> $OUT/ref_n$N.err; echo "ref rc=$?"; tail -c 300 $OUT/ref_n$N.json
