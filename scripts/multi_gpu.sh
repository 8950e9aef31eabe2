#!/bin/bash
# Multi-GPU pass on an N-GPU box: slab parity tests (torch.distributed ranks and the single-process cpml_multi
# handle, isotropic and viscoelastic, peer stores and NCCL), weak-scaling bench lines, and the compiled C++ driver
# with NGPU=N.  Usage: scripts/multi_gpu.sh N [steps]
N=$1; steps=${2:-60}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo$N.txt 2>&1
( time timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_drivers.py -m gpu -q -rs $PYTEST_EXTRA ) > gpurun_out/test_multi$N.log 2>&1
echo "pytest rc=$?" >> gpurun_out/test_multi$N.log
tail -8 gpurun_out/test_multi$N.log
: > gpurun_out/bench_n$N.jsonl
port=29510
for wl in ${WORKLOADS:-cfg3 cfg4 cfg5 cfg5d "cfg3 --halo sendrecv" "cfg5 --halo sendrecv"}; do
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
     bench.py --gpus $N --steps $steps --warmup 5 --workload $wl 2>> gpurun_out/bench_n$N.err | tail -1 >> gpurun_out/bench_n$N.jsonl
done
python - gpurun_out/bench_n$N.jsonl <<'PY'
import json, sys
for line in open(sys.argv[1]):
    try:
        d = json.loads(line)
        print(f"{d['config']['workload'][:60]:60s} n={d['n_gpus']} value {d['value']:.2f} Gpts/s step {d['ms_per_step']:.3f} ms frac {d['roofline']['frac']:.3f} e2e {d['e2e']['value']:.2f} launches {d['gpu_launches']} halo {str(d['run']['halo'])[:24]}")
    except Exception as e:
        print("FAILED", e, line[:200])
PY
make -C drivers -s
nz=$((640*N))
( time timeout 600 drivers/xseismic_cpml --program 3d_iso NZ=$nz NGPU=$N NSTEP=300 IT_DISPLAY=100 --no-images --out /tmp ) 2>&1 | grep -E "Gpts|z-slab|real" > gpurun_out/driver_n$N.txt
nzv=$((128*N))
( time timeout 600 drivers/xseismic_cpml --program 3d_visco NX=1024 NY=1024 NZ=$nzv NGPU=$N NSTEP=60 IT_DISPLAY=100 --no-images --out /tmp ) 2>&1 | grep -E "Gpts|z-slab|real" >> gpurun_out/driver_n$N.txt
cat gpurun_out/driver_n$N.txt
