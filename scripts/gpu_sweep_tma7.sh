#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/sweep_tma7.txt; : > $OUT
fmt='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d["roofline"]; li = d["config"]["launch"]
        print("  value %.2f Gpts/s  step %.3f ms  stress %.3f ms (%.3f)  vel %.3f ms (%.3f)  stepfrac %.3f  zc %d ctas %d st %d" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"], r["velocity_kernel"]["avg_launch_ms"], r["velocity_kernel"]["frac"], r["step"]["frac"], li["z_chunks"], li["ctas_stress"], li["stages"]))
    except Exception as e:
        print("  ?", l.strip()[:300])
'
run() { wl=$1; shift; echo "$wl $*" >> $OUT; env "$@" timeout 300 python bench.py --workload $wl --steps 40 --warmup 3 --no-cpu-baseline 2>&1 | python -c "$fmt" >> $OUT; }
run cfg3 CPML_STAGES=2
run cfg3 CPML_TY=4 CPML_MINB=2
run cfg3 CPML_ZCHUNKS=7
run cfg3 CPML_ZCHUNKS=8
run cfg3 CPML_ZCHUNKS=12
run cfg3 CPML_ZCHUNKS=18
run cfg4 CPML_TX=104
run cfg4 CPML_TX=104 CPML_ZCHUNKS=2
echo finished >> $OUT
