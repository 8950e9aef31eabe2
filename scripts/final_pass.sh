#!/bin/bash
# Final pass of a round on one B200: every GPU test, smoke, the default bench line and the reference arm the way the
# driver runs them, every workload, the single-precision tolerance numbers, compute-sanitizer memcheck over one case
# per kernel family, ncu summaries (stress + velocity kernel of one step) and the launch list of the final kernels.
tag=${1:-final}
mkdir -p gpurun_out
bash scripts/gpu_tests.sh $tag
timeout 600 python -m pytest tests/test_gpu_f32.py -m gpu -q -rP -k tolerance 2>&1 | grep "FP32 vs" > gpurun_out/f32_tolerance_$tag.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${tag}_default.json 2> gpurun_out/bench_${tag}_default.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_${tag}_reference.json 2> gpurun_out/bench_${tag}_reference.err
bash scripts/bench_variants.sh $tag 100 cfg3 cfg3f cfg4 cfg2 cfg6 cfg5 cfg5d $EXTRA_VARIANTS
bash scripts/sanitize.sh $tag > gpurun_out/sanitize_$tag.txt 2>&1
for wl in ${NCU_WORKLOADS:-cfg3 cfg4 cfg5 cfg2}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_(v?stress|v?velocity)" -s 8 -c 2 \
     -o gpurun_out/ncu_${tag}_$wl -f python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${tag}_$wl.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${tag}_cfg3.csv \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/launches_${tag}_cfg3.log 2>&1
