#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "2d" ) > gpurun_out/test_div4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_div4.log
timeout 300 python bench.py --workload cfg2 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/b_div4_cfg2.json 2> gpurun_out/b_div4_cfg2.err
