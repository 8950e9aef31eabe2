#!/bin/bash
# One `ncu --set full` capture of the update kernels of a workload (2 launches after the warm-up) and the launch list.
# Usage: scripts/ncu_capture.sh <tag> <workload> <kernel regex> [ENV=VALUE...]
tag=$1; wl=$2; rx=$3; shift 3
mkdir -p gpurun_out
env "$@" timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 8 -c 2 \
   -o gpurun_out/ncu_$tag -f python bench.py --workload "$wl" --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1
env "$@" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$tag.csv \
   python bench.py --workload "$wl" --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/launches_$tag.log 2>&1
