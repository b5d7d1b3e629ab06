#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): launch list of one bench step + one full ncu capture of the
# dominant kernel.  usage: scripts/gpu_profile.sh <tag>   -> gpurun_out/launches_<tag>.csv, prof_<tag>.ncu-rep
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c2 --no-atomic-peak"
# every launch of the bench (warm-up included; the summariser keeps the last complete build)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    $BENCH > gpurun_out/launches_${TAG}.stdout 2> gpurun_out/launches_${TAG}.stderr
# the dominant kernel, 4th launch (tables warm, sized from the previous build)
ncu --set full --clock-control none --import-source on -k regex:k_insert_windows -s 3 -c 1 \
    -o gpurun_out/prof_${TAG} -f $BENCH > gpurun_out/prof_${TAG}.stdout 2>&1
ls -la gpurun_out
