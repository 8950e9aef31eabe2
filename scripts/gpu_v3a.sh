#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_analytical_visco3d.py -x -q -m gpu -s ) > gpurun_out/test_v3a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_v3a.log
