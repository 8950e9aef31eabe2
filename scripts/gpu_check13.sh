#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_visco2d.py tests/test_gpu_parity.py tests/test_drivers.py -x -q -k "2d or cpp" ) > gpurun_out/test_gpu13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu13.log
( CPML_2D_PAIR=0 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "2d" ) > gpurun_out/test_gpu13_single.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu13_single.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke13.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke13.log
timeout 300 python bench.py --workload cfg2 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/b13_cfg2.json 2> gpurun_out/b13_cfg2.err
CPML_2D_PAIR=0 timeout 300 python bench.py --workload cfg2 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/b13_cfg2_single.json 2> gpurun_out/b13_cfg2_single.err
