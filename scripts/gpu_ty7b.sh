#!/bin/bash
# isotropic TMA kernels: per-kernel tiles (stress 104 x 7, velocity 104 x 8), finest-within-5% z chunks
mkdir -p gpurun_out
OUT=gpurun_out/sweep_ty7b.txt; : > $OUT
fmt='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d["roofline"]; c = d["config"]["launch"]
        print("  value %.2f Gpts/s  step %.3f ms  stress %.3f ms (%.3f)  vel %.3f ms (%.3f)  stepfrac %.3f e2e %.2f  vel tile %dx%d zc %d | stress ty %d zc %d" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"], r["velocity_kernel"]["avg_launch_ms"], r["velocity_kernel"]["frac"], r["step"]["frac"], d["e2e"]["value"], c["tile_x"], c["tile_y"], c["z_chunks"], c["stress_tile_y"], c["stress_z_chunks"]))
    except Exception as e:
        print("  ?", l.strip()[:300])
'
run() { wl=$1; shift; echo "$wl $*" >> $OUT; env "$@" timeout 300 python bench.py --workload $wl --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | python -c "$fmt" >> $OUT; }
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "3d" ) > gpurun_out/test_ty7b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_ty7b.log
run cfg3 DEFAULT=1
run cfg3 CPML_ZCHUNKS=9 CPML_ZCHUNKS_STRESS=16
run cfg3 CPML_ZCHUNKS=18 CPML_ZCHUNKS_STRESS=20
run cfg3 CPML_ZCHUNKS=18 CPML_ZCHUNKS_STRESS=32
run cfg3 CPML_TY_STRESS=8
run cfg4 DEFAULT=1
run cfg4 CPML_ZCHUNKS=1 CPML_ZCHUNKS_STRESS=1
echo finished >> $OUT
