# sweep of the centroid-range split of centroid_scores (engine reads PLAID_CSPLIT): WL=cfg2 scripts/csplit_sweep.sh 1 2 4 8
WL=${WL:-cfg4shard}
for cs in "$@"; do
  PLAID_CSPLIT=$cs timeout 600 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cs=$cs', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if v['ms_per_step']>0.2})"
done
