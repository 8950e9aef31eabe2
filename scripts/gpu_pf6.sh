#!/bin/bash
# viscoelastic kernels: L2 prefetch of the C-PML memory variables of the next plane (CPML_VPF bit 2)
mkdir -p gpurun_out
OUT=gpurun_out/sweep_vpf6.txt; : > $OUT
fmt='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d["roofline"]
        print("  value %.2f Gpts/s  step %.3f ms  stress %.3f ms (%.0f GB/s, %.3f)  vel %.3f ms (%.0f GB/s, %.3f)  stepfrac %.3f e2e %.2f" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["achieved"], r["frac"], r["velocity_kernel"]["avg_launch_ms"], r["velocity_kernel"]["achieved"], r["velocity_kernel"]["frac"], r["step"]["frac"], d["e2e"]["value"]))
    except Exception as e:
        print("  ?", l.strip()[:300])
'
run() { wl=$1; shift; echo "$wl $*" >> $OUT; env "$@" timeout 300 python bench.py --workload $wl --steps 30 --warmup 3 --no-cpu-baseline 2>&1 | python -c "$fmt" >> $OUT; }
( timeout 600 python -m pytest tests/test_gpu_visco.py -x -q ) > gpurun_out/test_pf6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_pf6.log
for wl in cfg5d cfg5; do run $wl CPML_VPF=2; run $wl CPML_VPF=6; done
echo finished >> $OUT
