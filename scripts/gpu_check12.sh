#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_visco2d.py tests/test_gpu_parity.py -x -q -k "2d" ) > gpurun_out/test_gpu12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu12.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke12.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke12.log
for wl in cfg2 cfg6; do
  timeout 300 python bench.py --workload $wl --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/b12_$wl.json 2> gpurun_out/b12_$wl.err
done
