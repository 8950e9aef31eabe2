#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/stream_mix > gpurun_out/stream_mix.txt 2>&1
OUT=gpurun_out/sweep_tma2.txt
: > $OUT
fmt='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d["roofline"]; li = d["config"]["launch"]
        print("  value %.2f Gpts/s  step %.3f ms  stress %.3f ms (%.0f GB/s, %.3f)  vel %.3f ms (%.0f GB/s, %.3f)  stepfrac %.3f e2e %.2f launches %d  %s" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["achieved"], r["frac"], r["velocity_kernel"]["avg_launch_ms"], r["velocity_kernel"]["achieved"], r["velocity_kernel"]["frac"], r["step"]["frac"], d["e2e"]["value"], d["gpu_launches"], li))
    except Exception as e:
        print("  ?", l.strip()[:300])
'
run() { wl=$1; shift; echo "$wl $*" >> $OUT; env "$@" timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | python -c "$fmt" >> $OUT; }
for wl in cfg3 cfg4; do
  run $wl CPML_TX=104 CPML_TY=4 CPML_STAGES=3 CPML_MINB=2
  run $wl CPML_TX=64 CPML_TY=8 CPML_STAGES=3 CPML_MINB=2
  run $wl CPML_TX=64 CPML_TY=4 CPML_STAGES=3 CPML_MINB=3
  run $wl CPML_TX=128 CPML_TY=2 CPML_STAGES=3 CPML_MINB=3
  run $wl CPML_TX=128 CPML_TY=2 CPML_STAGES=3 CPML_MINB=2 CPML_ZCHUNKS=20
  run $wl CPML_TX=128 CPML_TY=2 CPML_STAGES=3 CPML_MINB=2 CPML_ZCHUNKS=5
done
CPML_TX=104 CPML_TY=4 CPML_STAGES=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stress3d|k_velocity3d' -s 8 -c 2 \
   -o gpurun_out/prof_tma_104x4 -f python bench.py --workload cfg3 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tma1.log 2>&1
CPML_TX=128 CPML_TY=2 CPML_STAGES=3 CPML_MINB=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stress3d|k_velocity3d' -s 8 -c 2 \
   -o gpurun_out/prof_tma_128x2 -f python bench.py --workload cfg3 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tma2.log 2>&1
echo finished >> $OUT
