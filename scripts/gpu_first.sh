#!/bin/bash
# first GPU pass: parity tests, smoke, bench, tile sweep, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
for cfg in "32 8" "32 4" "32 16" "64 4" "64 2" "128 2" "128 1" "16 16" "16 8"; do
  set -- $cfg
  for zc in 0 2 8; do
    echo "TX=$1 TY=$2 ZCH=$zc" >> gpurun_out/sweep_cfg3.txt
    CPML_TX=$1 CPML_TY=$2 CPML_ZCHUNKS=$zc timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d['roofline']
        print('  value %.2f Gpts/s  step %.3f ms  stress %.3f ms (%.0f GB/s, %.3f)  vel %.3f ms (%.0f GB/s, %.3f)  e2e %.2f' % (d['value'], d['ms_per_step'], r['avg_launch_ms'], r['achieved'], r['frac'], r['velocity_kernel']['avg_launch_ms'], r['velocity_kernel']['achieved'], r['velocity_kernel']['frac'], d['e2e']['value']))
    except Exception as e:
        print('  ?', l.strip()[:300])
" >> gpurun_out/sweep_cfg3.txt
  done
done
timeout 600 python bench.py --workload cfg4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
timeout 600 python bench.py --workload cfg2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg3.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo finished > gpurun_out/done.txt
