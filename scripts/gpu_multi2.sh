#!/bin/bash
# two GPUs: slab tests (peer stores and NCCL send/recv) and the weak-scaling bench in both halo modes
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/test_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_multi.log
for wl in cfg3 cfg4; do
for halo in p2p sendrecv; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --workload $wl --halo $halo > gpurun_out/bench2_${wl}_${halo}.json 2> gpurun_out/bench2_${wl}_${halo}.err
done
timeout 300 python bench.py --steps 100 --warmup 5 --workload $wl --no-cpu-baseline > gpurun_out/bench1_${wl}.json 2> gpurun_out/bench1_${wl}.err
done
echo finished > gpurun_out/done2.txt
