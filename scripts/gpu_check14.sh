#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --workload cfg2 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/b14_cfg2_m2.json 2> gpurun_out/b14_cfg2_m2.err
CPML_2D_MINB=3 timeout 300 python bench.py --workload cfg2 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/b14_cfg2_m3.json 2> gpurun_out/b14_cfg2_m3.err
( CPML_2D_MINB=3 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "2d" ) > gpurun_out/test_gpu14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu14.log
