#!/bin/bash
# 2-GPU pass after the viscoelastic L2 prefetch: slab parity tests, weak-scaling bench lines
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/test_multi2e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_multi2e.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n2e_cfg3.json 2> gpurun_out/bench_n2e_cfg3.err; echo "rc=$?" >> gpurun_out/bench_n2e_cfg3.err
timeout 300 $TR bench.py --gpus 2 --workload cfg5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2e_cfg5.json 2> gpurun_out/bench_n2e_cfg5.err; echo "rc=$?" >> gpurun_out/bench_n2e_cfg5.err
timeout 300 $TR bench.py --gpus 2 --workload cfg5d --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2e_cfg5d.json 2> gpurun_out/bench_n2e_cfg5d.err; echo "rc=$?" >> gpurun_out/bench_n2e_cfg5d.err
