#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_visco.py -x -q ) > gpurun_out/test_visco.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_visco.log
