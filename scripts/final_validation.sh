#!/bin/bash
# One-GPU validation pass: GPU tests, bench lines of every workload, reference arm, ncu launch list and full capture, smoke.
# usage: scripts/final_validation.sh TAG   (files land in gpurun_out/TAG_*)
T=${1:-val}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; tail -2 gpurun_out/${T}_tests.log
timeout 900 python bench.py > gpurun_out/${T}_bench_cfg2_n1.json 2> gpurun_out/${T}_bench_cfg2_n1.err || echo "FAILED default bench"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err || echo "FAILED reference arm"
for wl in cfg1 cfg3 cfg5 codec; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-cfg4 > gpurun_out/${T}_bench_$wl.json 2> gpurun_out/${T}_bench_$wl.err || echo "FAILED $wl"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cfg4 > gpurun_out/${T}_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"centroid_scores|maxsim_fused|approx_scores|ivf_scores|ivf_pairs|select_top|compact_candidates|mark_candidates" -s 22 -c 22 -f -o gpurun_out/${T}_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cfg4 > gpurun_out/${T}_ncu_full.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python - "$T" <<'PY'
import json, sys
t = sys.argv[1]
for n in ["bench_cfg2_n1", "bench_reference", "bench_cfg1", "bench_cfg3", "bench_cfg5", "bench_codec"]:
    try:
        d = json.loads(open(f"gpurun_out/{t}_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d.get("ms_per_step", 0), 3), "value", d.get("value"), "e2e", d.get("e2e", {}).get("ms_per_step"), "roofline", d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(n, "unreadable", e)
PY
