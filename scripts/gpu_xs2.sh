#!/bin/bash
# velocity kernel: compile-time shell-first thread order (XS) on the default grid
mkdir -p gpurun_out
OUT=gpurun_out/sweep_xs2.txt; : > $OUT
fmt='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d["roofline"]
        print("  value %.2f Gpts/s  step %.3f ms  stress %.3f ms (%.3f)  vel %.3f ms (%.3f)  stepfrac %.3f e2e %.2f" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"], r["velocity_kernel"]["avg_launch_ms"], r["velocity_kernel"]["frac"], r["step"]["frac"], d["e2e"]["value"]))
    except Exception as e:
        print("  ?", l.strip()[:300])
'
run() { wl=$1; shift; echo "$wl $*" >> $OUT; env "$@" timeout 300 python bench.py --workload $wl --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | python -c "$fmt" >> $OUT; }
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "3d" ) > gpurun_out/test_xs2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_xs2.log
run cfg3 CPML_XSPLIT=0
run cfg3 CPML_XSPLIT=1
run cfg3 CPML_XSPLIT=0
run cfg3 CPML_XSPLIT=1
echo finished >> $OUT
