#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/tma_probe.log
for box in "32 8" "34 9" "106 5" "130 5"; do
for c in "0 0 1" "1 0 0" "2 0 0" "3 5 1" "-1 0 0" "-2 0 0" "0 -1 0" "-1 -1 0" "-2 -2 0" "30 40 9" "36 44 9" "0 0 -1" "0 0 10" "40 0 0" "-50 0 0"; do
  timeout 60 ./tools/tma_probe $box $c >> gpurun_out/tma_probe.log 2>&1
done; done
