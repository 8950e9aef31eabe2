#!/bin/bash
# viscoelastic kernels: L2 prefetch modes (CPML_VPF) + the analytical-solution test at the reference's full size
mkdir -p gpurun_out
OUT=gpurun_out/sweep_vpf.txt; : > $OUT
fmt='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); r = d["roofline"]
        print("  value %.2f Gpts/s  step %.3f ms  stress %.3f ms (%.0f GB/s, %.3f)  vel %.3f ms (%.0f GB/s, %.3f)  stepfrac %.3f e2e %.2f" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["achieved"], r["frac"], r["velocity_kernel"]["avg_launch_ms"], r["velocity_kernel"]["achieved"], r["velocity_kernel"]["frac"], r["step"]["frac"], d["e2e"]["value"]))
    except Exception as e:
        print("  ?", l.strip()[:300])
'
run() { wl=$1; shift; echo "$wl $*" >> $OUT; env "$@" timeout 300 python bench.py --workload $wl --steps 12 --warmup 3 --no-cpu-baseline 2>&1 | python -c "$fmt" >> $OUT; }
for wl in cfg5d cfg5; do
  for pf in 0 1 2 3 4; do run $wl CPML_VPF=$pf; done
done
( CPML_VPF=2 timeout 600 python -m pytest tests/test_gpu_visco.py -x -q ) > gpurun_out/test_vpf2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_vpf2.log
( CPML_VPF=4 timeout 600 python -m pytest tests/test_gpu_visco.py -x -q ) > gpurun_out/test_vpf4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_vpf4.log
( timeout 600 python -m pytest tests/test_analytical_visco2d.py tests/test_gpu_visco2d.py -x -q -m gpu -s ) > gpurun_out/test_analytic.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_analytic.log
echo finished >> $OUT
