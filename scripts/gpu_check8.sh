#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/test_gpu8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu8.log
for wl in cfg5d cfg5 cfg2 cfg6; do
  timeout 300 python bench.py --workload $wl --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/b8_$wl.json 2> gpurun_out/b8_$wl.err
done
