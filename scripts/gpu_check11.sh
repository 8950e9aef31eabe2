#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_visco.py tests/test_gpu_visco2d.py tests/test_gpu_parity.py -x -q -k "visco or 2d" ) > gpurun_out/test_gpu11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu11.log
for wl in cfg2 cfg5d cfg5 cfg6; do
  timeout 300 python bench.py --workload $wl --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/b11_$wl.json 2> gpurun_out/b11_$wl.err
done
