#!/bin/bash
# A/B bench runs on one box.  Usage: scripts/bench_variants.sh <tag> <steps> <variant>...
# where a variant is "workload[,ENV=VALUE...]", e.g. "cfg3,CPML_KERNEL=tma" ; one JSON line per variant in
# gpurun_out/bench_<tag>.jsonl plus a one-line summary each in gpurun_out/bench_<tag>.txt
tag=$1; steps=$2; shift 2
mkdir -p gpurun_out
: > gpurun_out/bench_$tag.jsonl; : > gpurun_out/bench_$tag.txt
for v in "$@"; do
  IFS=',' read -r -a parts <<< "$v"
  wl=${parts[0]}
  envs=("${parts[@]:1}")
  line=$(env "${envs[@]}" timeout 900 python bench.py --workload "$wl" --steps "$steps" --warmup 5 --no-cpu-baseline 2>> gpurun_out/bench_$tag.err | tail -1)
  echo "$line" >> gpurun_out/bench_$tag.jsonl
  python - "$v" "$line" >> gpurun_out/bench_$tag.txt <<'PY'
import json, sys
v, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    r = d["roofline"]; ks = r["kernels"]
    parts = [f"{k} {x['avg_launch_ms']:.3f} ms ({x['frac']:.3f})" for k, x in ks.items()]
    print(f"{v}\n  value {d['value']:.2f} Gpts/s  step {d['ms_per_step']:.3f} ms  stepfrac {r['frac']:.3f}  " + "  ".join(parts) +
          f"  e2e {d['e2e']['value']:.2f}  launches {d['gpu_launches']}")
except Exception as e:
    print(f"{v}\n  FAILED: {e}: {line[:200]}")
PY
done
cat gpurun_out/bench_$tag.txt
