#!/bin/bash
# 8-GPU pass: slab parity at world 8 (isotropic peer stores, viscoelastic NCCL), then the weak-scaling benches
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "(8 and p2p) or (visco and 8 and 4)" ) > gpurun_out/test_multi8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_multi8.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515"
timeout 300 $TR bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/bench_n8_cfg3.json 2> gpurun_out/bench_n8_cfg3.err; echo "rc=$?" >> gpurun_out/bench_n8_cfg3.err
timeout 300 $TR bench.py --gpus 8 --steps 40 --warmup 3 --workload cfg4 > gpurun_out/bench_n8_cfg4.json 2> gpurun_out/bench_n8_cfg4.err; echo "rc=$?" >> gpurun_out/bench_n8_cfg4.err
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 3 --workload cfg5 > gpurun_out/bench_n8_cfg5.json 2> gpurun_out/bench_n8_cfg5.err; echo "rc=$?" >> gpurun_out/bench_n8_cfg5.err
timeout 300 $TR bench.py --gpus 8 --steps 40 --warmup 3 --halo sendrecv > gpurun_out/bench_n8_cfg3_sendrecv.json 2> gpurun_out/bench_n8_cfg3_sendrecv.err; echo "rc=$?" >> gpurun_out/bench_n8_cfg3_sendrecv.err
echo finished > gpurun_out/done_multi8.txt
