#!/bin/bash
# Every GPU test + smoke on the box.  Usage: scripts/gpu_tests.sh [tag] [extra pytest args]
tag=${1:-run}; shift
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q "$@" ) > gpurun_out/test_gpu_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu_$tag.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$tag.log
tail -5 gpurun_out/test_gpu_$tag.log; tail -2 gpurun_out/smoke_$tag.log
