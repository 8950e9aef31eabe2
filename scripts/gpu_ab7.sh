#!/bin/bash
# A/B on one box: HEAD library (ab/libcpml_b200_head.so) against the working tree's
mkdir -p gpurun_out
OUT=gpurun_out/sweep_ab7.txt; : > $OUT
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
( timeout 600 python -m pytest tests/test_gpu_visco.py -x -q ) > gpurun_out/test_ab7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_ab7.log
cp seismic_cpml_b200/libcpml_b200.so /tmp/new.so
for wl in cfg5d cfg5; do
  cp ab/libcpml_b200_head.so seismic_cpml_b200/libcpml_b200.so; run $wl LIB=head CPML_VPF=2
  cp /tmp/new.so seismic_cpml_b200/libcpml_b200.so; run $wl LIB=new CPML_VPF=2; run $wl LIB=new CPML_VPF=6
done
echo finished >> $OUT
