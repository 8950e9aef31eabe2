#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_visco.py tests/test_gpu_multi.py -x -q -m gpu ) > gpurun_out/test_szz.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_szz.log
timeout 300 python bench.py --workload cfg5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b_szz_cfg5.json 2> gpurun_out/b_szz_cfg5.err
