#!/bin/bash
# viscoelastic kernel variants: cp.async staging on/off x register caps x tiles
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_visco.py tests/test_gpu_parity.py -x -q -k "visco or 2d" ) > gpurun_out/test_gpu_v4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu_v4.log
OUT=gpurun_out/sweep_visco3.txt; : > $OUT
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
  for spec in "32 8 1 0" "32 8 1 1" "32 8 2 0" "32 8 2 1" "32 4 2 0" "32 4 2 1" "32 4 3 0" "32 4 3 1" "32 4 4 0" "32 4 4 1" "64 2 2 1" "64 2 4 1"; do
    set -- $spec
    run $wl CPML_VTX=$1 CPML_VTY=$2 CPML_VMINB=$3 CPML_VASYNC=$4
  done
done
echo finished > gpurun_out/done_visco4.txt
