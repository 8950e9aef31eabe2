#!/bin/bash
# viscoelastic kernels at 12 warps per SM (32 x 4 tile, 3 resident blocks: 168-register cap) against the default 16
mkdir -p gpurun_out
OUT=gpurun_out/sweep_v12w.txt; : > $OUT
fmt='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d["roofline"]
        print("  value %.2f Gpts/s  step %.3f ms  stress %.3f ms (%.3f)  vel %.3f ms (%.3f)  stepfrac %.3f" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"], r["velocity_kernel"]["avg_launch_ms"], r["velocity_kernel"]["frac"], r["step"]["frac"]))
    except Exception as e:
        print("  ?", l.strip()[:300])
'
run() { wl=$1; shift; echo "$wl $*" >> $OUT; env "$@" timeout 300 python bench.py --workload $wl --steps 30 --warmup 3 --no-cpu-baseline 2>&1 | python -c "$fmt" >> $OUT; }
for wl in cfg5d cfg5; do
  run $wl DEFAULT=1
  run $wl CPML_VTX=32 CPML_VTY=4 CPML_VMINB=3
  run $wl CPML_VTX=32 CPML_VTY=4 CPML_VMINB=4
  run $wl CPML_VKCHUNK=32
  run $wl CPML_VKCHUNK=8
done
echo finished >> $OUT
