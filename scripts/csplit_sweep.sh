for cs in 3 6 10; do
  PLAID_CSPLIT=$cs timeout 600 python bench.py --workload cfg4shard --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cs=$cs', round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items() if v['ms_per_step']>0.3})"
done
