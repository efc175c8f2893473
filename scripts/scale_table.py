#!/usr/bin/env python
"""profiles/rNN_scale_cfg4.md from the bench lines of a 1/2/4/8-GPU sweep: usage scale_table.py n1.json n2.json n4.json n8.json"""
import json
import sys

rows = [json.load(open(p)) for p in sys.argv[1:]]
base2, base4 = rows[0], rows[0]["cfg4"]
print("# 1 / 2 / 4 / 8 B200: cfg2 per rank (weak) and the fixed 10M-passage cfg4 collection (strong), round 2 final code\n")
print("`python -m torch.distributed.run --nproc-per-node N bench.py --gpus N --steps 20 --warmup 5` (N = 1: `python bench.py`), one")
print("fresh box per N; device-timed with CUDA events, max over ranks.  cfg2: one 112k-passage shard PER RANK (the collection grows")
print("with N), efficiency = ms(1) / ms(N).  cfg4: ONE 10M-passage collection pid-sharded over the ranks, 1024 queries,")
print("efficiency = queries/s(N) / (N x queries/s(1)).\n")
print("| N | cfg2 ms/step | cfg2 queries/s (x N shards) | weak eff. | cfg2 e2e ms (plugin call) | e2e weak eff. | cfg4 per_shard ms | queries/s | strong eff. | cfg4 exact ms | queries/s | strong eff. | cfg4 e2e per_shard / exact ms |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for d in rows:
    n = d["n_gpus"]
    c = d["cfg4"]
    x = c.get("exact_global") or {"ms_per_step": c["ms_per_step"], "queries_per_s": c["queries_per_s"], "e2e": c["e2e"]}
    print(f"| {n} | {d['ms_per_step']:.2f} | {d['queries_per_s'] / 1e3:.1f} k | {base2['ms_per_step'] / d['ms_per_step']:.3f} | "
          f"{d['e2e']['ms_per_step']:.2f} | {base2['e2e']['ms_per_step'] / d['e2e']['ms_per_step']:.3f} | "
          f"{c['ms_per_step']:.2f} | {c['queries_per_s'] / 1e3:.1f} k | {c['queries_per_s'] / (n * base4['queries_per_s']):.2f} | "
          f"{x['ms_per_step']:.2f} | {x['queries_per_s'] / 1e3:.1f} k | {x['queries_per_s'] / (n * base4['queries_per_s']):.2f} | "
          f"{c['e2e']['ms_per_step']:.2f} / {x['e2e']['ms_per_step']:.2f} |")
print("\nPer-stage device time on rank 0, cfg4 (ms per 1024 queries; per_shard | exact):\n")
stages = ["centroid_scores", "candidates", "filter_stage1", "select1", "exchange_stage1", "filter_stage2", "select2", "exchange_stage2",
          "maxsim_fused", "topk"]
print("| N | " + " | ".join(stages) + " |")
print("|---|" + "---|" * len(stages))
for d in rows:
    c = d["cfg4"]
    k1 = {s: v["ms_per_step"] for s, v in c["kernels_rank0"].items()}
    k2 = (c.get("exact_global") or {}).get("kernels_rank0", {})
    print(f"| {d['n_gpus']} | " + " | ".join(f"{k1.get(s, 0):.2f} \\| {k2[s]:.2f}" if s in k2 else f"{k1.get(s, 0):.2f}" for s in stages) + " |")
for d in rows:
    c = d["cfg4"]
    print(f"\nN={d['n_gpus']}: {c['queries_per_chunk']} queries per chunk, {c['passages_per_rank']} passages / {c['index_bytes_per_rank'] / 1e9:.1f} GB per rank, "
          f"T1/T2/T3 tokens per query over all shards {c['T1_tokens_per_query_all_shards'] / 1e6:.1f} M / {c['T2_tokens_per_query_all_shards'] / 1e3:.0f} k / "
          f"{c['T3_tokens_per_query_all_shards'] / 1e3:.0f} k, stage-1 scan fallback for {int(c['stage1_scan_fallback_queries_all_shards'])} (query, shard) pairs, "
          f"centroid_scores at {c['kernels_rank0']['centroid_scores'].get('frac_hbm')} of HBM; clocks {d['clocks']}")
