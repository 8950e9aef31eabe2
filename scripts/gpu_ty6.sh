#!/bin/bash
# isotropic TMA kernels on 1024-wide slabs: 128 x 6 tile (384 threads, 168-register cap) against 128 x 8
mkdir -p gpurun_out
OUT=gpurun_out/sweep_ty6.txt; : > $OUT
fmt='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d["roofline"]; c = d["config"]["launch"]
        print("  value %.2f Gpts/s  step %.3f ms  stress %.3f ms (%.3f)  vel %.3f ms (%.3f)  stepfrac %.3f e2e %.2f  vel tile %dx%d zc %d | stress ty %d zc %d" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"], r["velocity_kernel"]["avg_launch_ms"], r["velocity_kernel"]["frac"], r["step"]["frac"], d["e2e"]["value"], c["tile_x"], c["tile_y"], c["z_chunks"], c["stress_tile_y"], c["stress_z_chunks"]))
    except Exception as e:
        print("  ?", l.strip()[:300])
'
run() { wl=$1; shift; echo "$wl $*" >> $OUT; env "$@" timeout 300 python bench.py --workload $wl --steps 40 --warmup 4 --no-cpu-baseline 2>&1 | python -c "$fmt" >> $OUT; }
( CPML_TX=128 CPML_TY=6 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "test_3d_iso_matches_oracle or slabs_with_peer" ) > gpurun_out/test_ty6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_ty6.log
run cfg4 DEFAULT=1
run cfg4 CPML_TY_STRESS=6
run cfg4 CPML_TY_STRESS=6 CPML_ZCHUNKS_STRESS=3
run cfg4 CPML_TY=6
run cfg4 CPML_TY_STRESS=6 CPML_ZCHUNKS_STRESS=1
echo finished >> $OUT
