#!/bin/bash
# isotropic TMA kernels: 104 x 7 tile on 384 threads (12 warps -> 168-register cap, no spills)
mkdir -p gpurun_out
OUT=gpurun_out/sweep_ty7.txt; : > $OUT
fmt='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d["roofline"]; c = d["config"]["launch"]
        print("  value %.2f Gpts/s  step %.3f ms  stress %.3f ms (%.3f)  vel %.3f ms (%.3f)  stepfrac %.3f e2e %.2f  tile %dx%d zc %d items %d" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"], r["velocity_kernel"]["avg_launch_ms"], r["velocity_kernel"]["frac"], r["step"]["frac"], d["e2e"]["value"], c["tile_x"], c["tile_y"], c["z_chunks"], c["items"]))
    except Exception as e:
        print("  ?", l.strip()[:300])
'
run() { wl=$1; shift; echo "$wl $*" >> $OUT; env "$@" timeout 300 python bench.py --workload $wl --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | python -c "$fmt" >> $OUT; }
( CPML_TX=104 CPML_TY=7 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "3d and not tma_tiles" ) > gpurun_out/test_ty7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_ty7.log
run cfg3 CPML_TY=8
run cfg3 CPML_TY=7
run cfg3 CPML_TY=7 CPML_ZCHUNKS=8
run cfg3 CPML_TY=7 CPML_ZCHUNKS=16
run cfg3 CPML_TY=7 CPML_ZCHUNKS=5
run cfg3 CPML_TY=7 CPML_STAGES=3
echo finished >> $OUT
