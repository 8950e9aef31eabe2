#!/bin/bash
# last pass of the round: every GPU test, smoke, the default bench line and the reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/test_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu_final.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_final.log
timeout 600 python bench.py > gpurun_out/bench_final_cfg3.json 2> gpurun_out/bench_final_cfg3.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err
