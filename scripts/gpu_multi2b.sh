#!/bin/bash
# 2-GPU pass: all multi-GPU parity tests (isotropic p2p / sendrecv, viscoelastic), then N=2 benches
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/test_multi2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_multi2b.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout 300 $TR bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/bench_n2_cfg3.json 2> gpurun_out/bench_n2_cfg3.err; echo "rc=$?" >> gpurun_out/bench_n2_cfg3.err
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --workload cfg5 > gpurun_out/bench_n2_cfg5.json 2> gpurun_out/bench_n2_cfg5.err; echo "rc=$?" >> gpurun_out/bench_n2_cfg5.err
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --workload cfg5d > gpurun_out/bench_n2_cfg5d.json 2> gpurun_out/bench_n2_cfg5d.err; echo "rc=$?" >> gpurun_out/bench_n2_cfg5d.err
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 4 --warmup 1 > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err; echo "rc=$?" >> gpurun_out/bench_n2_ref.err
echo finished > gpurun_out/done_multi2b.txt
