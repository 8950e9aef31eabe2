#!/bin/bash
mkdir -p gpurun_out
CPML_TX=104 CPML_TY=8 CPML_MINB=1 CPML_STAGES=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stress3d|k_velocity3d' -s 8 -c 2 \
   -o gpurun_out/prof_pairs_104x8 -f python bench.py --workload cfg3 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_pairs.log 2>&1
