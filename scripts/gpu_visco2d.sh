#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_visco2d.py tests/test_abi_host.py -x -q ) > gpurun_out/test_visco2d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_visco2d.log
timeout 600 python bench.py --workload cfg6 --steps 100 --warmup 5 > gpurun_out/bench_cfg6.json 2> gpurun_out/bench_cfg6.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_vstress2d|k_vvelocity2d' -s 8 -c 2 \
   -o gpurun_out/prof_cfg6 -f python bench.py --workload cfg6 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cfg6.log 2>&1
