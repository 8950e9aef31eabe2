#!/bin/bash
# 2-GPU pass: slab parity tests (p2p + sendrecv), then bench N=2 in both halo modes, N=1 for the same box
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/test_multi2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_multi2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --steps 40 --warmup 5 --halo sendrecv > gpurun_out/bench_n2_sendrecv.json 2> gpurun_out/bench_n2_sendrecv.err; echo "rc=$?" >> gpurun_out/bench_n2_sendrecv.err
timeout 300 $TR bench.py --gpus 2 --steps 40 --warmup 5 --halo p2p > gpurun_out/bench_n2_p2p.json 2> gpurun_out/bench_n2_p2p.err; echo "rc=$?" >> gpurun_out/bench_n2_p2p.err
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --workload cfg4 --halo p2p > gpurun_out/bench_n2_cfg4_p2p.json 2> gpurun_out/bench_n2_cfg4_p2p.err; echo "rc=$?" >> gpurun_out/bench_n2_cfg4_p2p.err
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --workload cfg4 --halo sendrecv > gpurun_out/bench_n2_cfg4_sendrecv.json 2> gpurun_out/bench_n2_cfg4_sendrecv.err; echo "rc=$?" >> gpurun_out/bench_n2_cfg4_sendrecv.err
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo finished > gpurun_out/done_multi2.txt
