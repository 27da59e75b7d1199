#!/bin/bash
# Round-2 GPU check: parity tests (all, no -x), default bench, host topology notes.
# gpurun --timeout 1500 -- 'bash tools/gpu_r2.sh <tag> [pytest -k expr]'
TAG=${1:-r2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
KEXPR=${2:-}
{ nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv; nvidia-smi topo -m; lscpu | head -30; ls /sys/devices/system/node/; \
  cat /sys/devices/system/node/node*/meminfo | grep MemTotal; nproc; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; } > $OUT/host.txt 2>&1
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -m gpu -q --timeout 240 -k "$KEXPR" > $OUT/pytest.log 2>&1
else
  timeout 900 python -m pytest tests -m gpu -q --timeout 240 > $OUT/pytest.log 2>&1
fi
echo "pytest rc=$?" >> $OUT/pytest.log
tail -40 $OUT/pytest.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 6000 $OUT/bench.json; tail -5 $OUT/bench.err
