#!/bin/bash
mkdir -p gpurun_out
for spec in "cfg3 16 8" "cfg4 32 8"; do
  set -- $spec
  CPML_TX=$2 CPML_TY=$3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stress3d|k_velocity3d' -s 8 -c 2 \
     -o gpurun_out/prof2_$1 -f python bench.py --workload $1 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2_$1.log 2>&1
done
