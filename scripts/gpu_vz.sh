#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_visco.py tests/test_drivers.py -x -q -m gpu -k "vz or driver or golden or matches_oracle" ) > gpurun_out/test_vz.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_vz.log
