#!/bin/bash
# full single-GPU pass (round-1 v9): all GPU tests, smoke, every bench workload, reference arms, launch lists, ncu full
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/test_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
timeout 300 python bench.py --workload cfg4 --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
timeout 300 python bench.py --workload cfg2 --steps 100 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
timeout 600 python bench.py --workload cfg5d --steps 40 --warmup 3 > gpurun_out/bench_cfg5d.json 2> gpurun_out/bench_cfg5d.err
timeout 600 python bench.py --workload cfg5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
timeout 600 python bench.py --workload cfg6 --steps 100 --warmup 5 > gpurun_out/bench_cfg6.json 2> gpurun_out/bench_cfg6.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg3.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stress3d|k_velocity3d' -s 8 -c 2 \
   -o gpurun_out/prof_cfg3_v9 -f python bench.py --workload cfg3 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cfg3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_vstress3d|k_vvelocity3d' -s 8 -c 2 \
   -o gpurun_out/prof_cfg5d_v9 -f python bench.py --workload cfg5d --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cfg5d.log 2>&1
echo finished > gpurun_out/done_full12.txt
