#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/sweep_kmod.txt; : > $OUT
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
( timeout 600 python -m pytest tests/test_gpu_visco.py -x -q ) > gpurun_out/test_kmod.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_kmod.log
run cfg5d DEFAULT=1
run cfg5 DEFAULT=1
echo finished >> $OUT
