#!/bin/bash
# after div_exact: all single-GPU parity tests, then visco + 2-D benches, tile sweep, ncu
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/test_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu.log
timeout 600 python bench.py --workload cfg5d --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5d.json 2> gpurun_out/bench_cfg5d.err
timeout 600 python bench.py --workload cfg5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
timeout 300 python bench.py --workload cfg2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
OUT=gpurun_out/sweep_visco2.txt; : > $OUT
fmt='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d["roofline"]
        print("  value %.2f Gpts/s  step %.3f ms  stress %.3f ms (%.0f GB/s, %.3f)  vel %.3f ms (%.0f GB/s, %.3f)  stepfrac %.3f e2e %.2f" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["achieved"], r["frac"], r["velocity_kernel"]["avg_launch_ms"], r["velocity_kernel"]["achieved"], r["velocity_kernel"]["frac"], r["step"]["frac"], d["e2e"]["value"]))
    except Exception as e:
        print("  ?", l.strip()[:300])
'
run() { wl=$1; shift; echo "$wl $*" >> $OUT; env "$@" timeout 300 python bench.py --workload $wl --steps 12 --warmup 3 --no-cpu-baseline 2>&1 | python -c "$fmt" >> $OUT; }
for wl in cfg5d cfg5; do
  for spec in "32 8 16" "32 4 16" "32 4 32" "64 2 16" "16 16 16" "32 4 8"; do
    set -- $spec
    run $wl CPML_VTX=$1 CPML_VTY=$2 CPML_VKCHUNK=$3
  done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_vstress3d|k_vvelocity3d' -s 8 -c 2 \
   -o gpurun_out/prof_cfg5d_v2 -f python bench.py --workload cfg5d --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cfg5d.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stress2d|k_velocity2d' -s 8 -c 2 \
   -o gpurun_out/prof_cfg2 -f python bench.py --workload cfg2 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cfg2.log 2>&1
echo finished > gpurun_out/done_visco3.txt
