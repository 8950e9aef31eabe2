#!/bin/bash
# tile / region sweep of the 3-D kernels (bench.py, roofline fields), results in gpurun_out/sweep2.txt
mkdir -p gpurun_out
OUT=gpurun_out/sweep2.txt
: > $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "3d" > gpurun_out/test_gpu3d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu3d.log
fmt='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d["roofline"]
        print("  value %.2f Gpts/s  step %.3f ms  stress %.3f ms (%.0f GB/s, %.3f)  vel %.3f ms (%.0f GB/s, %.3f)  stepfrac %.3f e2e %.2f launches %d" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["achieved"], r["frac"], r["velocity_kernel"]["avg_launch_ms"], r["velocity_kernel"]["achieved"], r["velocity_kernel"]["frac"], r["step"]["frac"], d["e2e"]["value"], d["gpu_launches"]))
    except Exception as e:
        print("  ?", l.strip()[:300])
'
run() { # workload, env...
  wl=$1; shift
  echo "$wl $*" >> $OUT
  env "$@" timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | python -c "$fmt" >> $OUT
}
for wl in cfg3 cfg4; do
  run $wl CPML_REGIONS=1
  for t in "32 8" "32 4" "16 8" "64 4" "128 2" "16 16" "64 2"; do
    set -- $t
    run $wl CPML_TX=$1 CPML_TY=$2
  done
  run $wl CPML_TX=32 CPML_TY=8 CPML_ZCHUNKS=4
  run $wl CPML_TX=32 CPML_TY=8 CPML_ZCHUNKS=32
  run $wl CPML_TX=32 CPML_TY=8 CPML_PML_TX=32 CPML_PML_TY=4
done
echo finished >> $OUT
