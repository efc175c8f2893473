set -x
mkdir -p gpurun_out
for wl in cfg2 cfg1 cfg3 cfg4shard cfg5; do
  extra=""
  [ "$wl" = "cfg2" ] || extra="--no-cpu-baseline"
  timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 $extra > gpurun_out/r01b_$wl.json 2> gpurun_out/r01b_$wl.err || echo "FAILED $wl"
  tail -c 600 gpurun_out/r01b_$wl.json | head -c 600; echo
done
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-fused --no-cpu-baseline > gpurun_out/r01b_cfg2_unfused.json 2> gpurun_out/r01b_cfg2_unfused.err || echo FAILED unfused
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2
