#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): one `ncu --set full` capture per "kernel:skip" argument over
# scripts/ncu_target.py (device-resident builds of the C5 shard).  usage: scripts/ncu_kernel.sh <tag> <kernel>:<skip> ...
set -u
TAG=$1; shift
mkdir -p gpurun_out
for KS in "$@"; do
  K=${KS%%:*}; S=${KS##*:}
  ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -o gpurun_out/prof_${TAG}_${K} -f \
      python scripts/ncu_target.py c5 1250000 5 3 > gpurun_out/prof_${TAG}_${K}.stdout 2>&1
  ncu -i gpurun_out/prof_${TAG}_${K}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_${K}.csv 2>/dev/null
done
ls gpurun_out | tail -5
