#!/bin/bash
# A/B runs of bench.py on one box: each argument is "label|extra bench args|env assignments"
# usage: scripts/ab_bench.sh "base||" "s2|--streams 2|" ...
mkdir -p gpurun_out
for spec in "$@"; do
  IFS='|' read -r label bargs envs <<< "$spec"
  env $envs python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg4 $bargs > gpurun_out/ab_$label.json 2> gpurun_out/ab_$label.err || tail -5 gpurun_out/ab_$label.err
  python - "$label" <<'PY'
import json, sys
lab = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/ab_{lab}.json"))
    ks = " ".join(f"{k}={v['ms_per_step']:.3f}" for k, v in d["kernels"].items())
    print(f"{lab}: step {d['ms_per_step']:.3f} ms, e2e {d['e2e']['ms_per_step']:.3f} | {ks}")
except Exception as e:
    print(lab, "failed", e)
PY
done
