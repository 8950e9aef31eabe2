#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "3d" ) > gpurun_out/test_gpu9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu9.log
for wl in cfg3 cfg4; do
  timeout 300 python bench.py --workload $wl --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/b9_$wl.json 2> gpurun_out/b9_$wl.err
  CPML_STAGES=3 timeout 300 python bench.py --workload $wl --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/b9_${wl}_st3.json 2> gpurun_out/b9_${wl}_st3.err
done
