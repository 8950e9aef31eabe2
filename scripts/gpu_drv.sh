#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_drivers.py tests/test_analytical_elastic.py -x -q -m gpu -s ) > gpurun_out/test_drv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_drv.log
