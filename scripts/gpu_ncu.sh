#!/bin/bash
# ncu --set full capture of the two hot kernels (cfg3 default tile), one GPU
mkdir -p gpurun_out
WL=${1:-cfg3}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stress3d|k_velocity3d' -s 8 -c 4 \
   -o gpurun_out/prof_${WL} -f python bench.py --workload $WL --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${WL}.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_${WL}.log
