#!/bin/bash
# A/B timing of kernel variants built with `python -m reranking_multimodal_retrievers_b200.build --variant NAME -DFLAGS`:
# runs the default bench once per library and prints the per-kernel ms of each.  Usage: scripts/ab_variants.sh NAME...
V=reranking_multimodal_retrievers_b200/csrc/build/variants
mkdir -p gpurun_out
for name in "$@"; do
  lib=""
  [ "$name" != "default" ] && lib="$PWD/$V/libplaid_b200.$name.so"
  PLAID_B200_LIB=$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} 2> gpurun_out/ab_$name.err > gpurun_out/ab_$name.json
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{n}.json").read().strip().splitlines()[-1])
    print(n, "ms/step", round(d["ms_per_step"], 3), {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items() if v["ms_per_step"] > 0.1})
except Exception as e:
    print(n, "failed", e)
PY
done
