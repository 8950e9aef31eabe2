#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/test_multi2c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_multi2c.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/bench_n2c_cfg3.json 2> gpurun_out/bench_n2c_cfg3.err; echo "rc=$?" >> gpurun_out/bench_n2c_cfg3.err
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 4 --warmup 1 > gpurun_out/bench_n2c_ref.json 2> gpurun_out/bench_n2c_ref.err
