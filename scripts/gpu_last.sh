#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/test_gpu_last.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu_last.log
