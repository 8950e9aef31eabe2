#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/sweep3.txt
: > $OUT
fmt='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d["roofline"]
        print("  value %.2f Gpts/s  step %.3f ms  stress %.3f ms (%.0f GB/s, %.3f)  vel %.3f ms (%.0f GB/s, %.3f)  stepfrac %.3f e2e %.2f launches %d" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["achieved"], r["frac"], r["velocity_kernel"]["avg_launch_ms"], r["velocity_kernel"]["achieved"], r["velocity_kernel"]["frac"], r["step"]["frac"], d["e2e"]["value"], d["gpu_launches"]))
    except Exception as e:
        print("  ?", l.strip()[:300])
'
run() { wl=$1; shift; echo "$wl $*" >> $OUT; env "$@" timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | python -c "$fmt" >> $OUT; }
for wl in cfg3 cfg4; do
  for t in "32 8" "32 4" "16 8" "64 4" "128 2" "16 16" "64 2" "128 1"; do
    set -- $t
    run $wl CPML_TX=$1 CPML_TY=$2
  done
  run $wl CPML_TX=32 CPML_TY=8 CPML_ZCHUNKS=8
  run $wl CPML_TX=32 CPML_TY=8 CPML_ZCHUNKS=40
done
for wl in cfg3 cfg4; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 42 -c 14 --csv --log-file gpurun_out/launches_regions_$wl.csv python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_regions_$wl.log 2>&1
done
echo finished >> $OUT
