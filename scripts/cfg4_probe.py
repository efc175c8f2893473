#!/usr/bin/env python
"""Development aid: the 10M-passage cfg4 collection (or its first --blocks blocks of 1.25M passages) on ONE GPU,
a few search passes over the 1024 bench queries -- the thing to run under `ncu --metrics gpu__time_duration.sum`
for a per-kernel launch list of the large-shard path.  usage: cfg4_probe.py [--blocks 8] [--passes 2]"""
import argparse
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from reranking_multimodal_retrievers_b200 import Searcher, synthetic  # noqa: E402
from reranking_multimodal_retrievers_b200.index import DeviceIndex  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--blocks", type=int, default=8)
ap.add_argument("--passes", type=int, default=2)
args = ap.parse_args()
c = bench.CFG4
dev = torch.device("cuda", 0)
shard = synthetic.make_collection_shard(range(args.blocks), c["passages_per_block"], c["lo"], c["hi"], c["nbits"], c["C"],
                                        c["blocks"], seed=4000, device=dev)
index = DeviceIndex(shard, dev)
ppb = c["passages_per_block"]
view = types.SimpleNamespace(codes=shard.codes, residuals=shard.residuals, doclens=shard.doclens[:ppb], centroids=shard.centroids,
                             bucket_weights=shard.bucket_weights, nbits=shard.nbits, dim=shard.dim, num_passages=ppb)
Q = synthetic.make_queries(view, c["B"], c["Lq"], seed=199).to(dev)
del shard
eng = Searcher(index=index).ranker.engine
for i in range(args.passes):
    eng.events = [] if i == args.passes - 1 else None
    out = eng.search_batch(Q, k=c["k"], remove_zero_rows=True)
    torch.cuda.synchronize()
eng.check_flags()
ms = {}
for stage, a, b in eng.events:
    ms[stage] = ms.get(stage, 0.0) + a.elapsed_time(b)
print({k: round(v, 3) for k, v in ms.items()}, "total", round(sum(ms.values()), 2))
