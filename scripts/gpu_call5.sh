#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "3d" ) > gpurun_out/test_gpu3d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu3d.log
grep -q "rc=0" gpurun_out/test_gpu3d.log || { echo "tests failed, sweep skipped" > gpurun_out/sweep_tma4.txt; exit 0; }
OUT=gpurun_out/sweep_tma4.txt
: > $OUT
fmt='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d["roofline"]; li = d["config"]["launch"]
        print("  value %.2f Gpts/s  step %.3f ms  stress %.3f ms (%.0f GB/s, %.3f)  vel %.3f ms (%.0f GB/s, %.3f)  stepfrac %.3f e2e %.2f  zc %d ctas %d st %d" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["achieved"], r["frac"], r["velocity_kernel"]["avg_launch_ms"], r["velocity_kernel"]["achieved"], r["velocity_kernel"]["frac"], r["step"]["frac"], d["e2e"]["value"], li["z_chunks"], li["ctas_stress"], li["stages"]))
    except Exception as e:
        print("  ?", l.strip()[:300])
'
run() { wl=$1; shift; echo "$wl $*" >> $OUT; env "$@" timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | python -c "$fmt" >> $OUT; }
for wl in cfg3 cfg4; do
  for spec in "128 7 1" "128 3 2" "128 8 1" "104 8 1"; do
    set -- $spec
    for st in 2 3; do
      run $wl CPML_TX=$1 CPML_TY=$2 CPML_MINB=$3 CPML_STAGES=$st
    done
  done
done
echo finished >> $OUT
