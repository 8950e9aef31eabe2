#!/bin/bash
# viscoelastic: bench on the default grid and on the 1024x1024x128 slab, tile sweep, launch list, ncu full
mkdir -p gpurun_out
timeout 600 python bench.py --workload cfg5d --steps 30 --warmup 3 > gpurun_out/bench_cfg5d.json 2> gpurun_out/bench_cfg5d.err
timeout 600 python bench.py --workload cfg5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
OUT=gpurun_out/sweep_visco1.txt; : > $OUT
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
  for spec in "32 8 16" "32 8 8" "32 8 32" "32 4 16" "64 4 16" "64 2 16" "128 2 16" "16 16 16"; do
    set -- $spec
    run $wl CPML_VTX=$1 CPML_VTY=$2 CPML_VKCHUNK=$3
  done
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg5d.csv python bench.py --workload cfg5d --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_v.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_vstress3d|k_vvelocity3d' -s 8 -c 2 \
   -o gpurun_out/prof_cfg5d -f python bench.py --workload cfg5d --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cfg5d.log 2>&1
echo finished > gpurun_out/done_visco2.txt
