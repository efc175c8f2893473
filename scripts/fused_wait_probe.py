#!/usr/bin/env python3
"""Development aid: per-warp share of the fused MaxSim kernel's decompressor time spent waiting for a free stage.
Needs a library built with -DMS_DBG_TIMING (python -m reranking_multimodal_retrievers_b200.build --variant dbgt -DMS_DBG_TIMING)
selected through PLAID_B200_LIB.  Prints, per decompressor warp index (0..15 = group * 4 + unit), the wait share averaged
over the CTAs of the last launch."""
import ctypes, sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reranking_multimodal_retrievers_b200 import _lib, synthetic
from reranking_multimodal_retrievers_b200.engine import SearchEngine
from reranking_multimodal_retrievers_b200.index import DeviceIndex

sx = synthetic.make_synthetic_index(112_000, 120, 239, 2, seed=1234, mode="codes", device="cuda")
Q = synthetic.make_queries(sx, 512, 64, seed=99).to(torch.float32)
eng = SearchEngine(DeviceIndex(sx))
for _ in range(3):
    eng.search_batch(Q, k=100)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (160 * 16 * 8))()
rc = _lib.lib().plaid_debug_read_fused_waits(buf)
a = np.frombuffer(buf, dtype=np.uint64).reshape(160, 16, 8).astype(np.float64)[:147]
share = a[:, :, 0] / np.maximum(a[:, :, 1], 1)
print("rc", rc, "mean total cycles", a[:, :, 1].mean())
print("wait share per decompressor warp (mean over CTAs), rows = groups, columns = unit in the tile:")
print(np.round(share.mean(axis=0).reshape(4, 4), 3))
print("per CTA mean: min %.3f max %.3f" % (share.mean(axis=1).min(), share.mean(axis=1).max()))
tot = a[:, :, 1].sum()
print("share of decompressor time: all waits %.3f, first unit of an item %.3f, waits > 3000 cycles %.3f" % (a[:, :, 0].sum() / tot, a[:, :, 2].sum() / tot, a[:, :, 3].sum() / tot))
print("units per warp %.0f, waits > 200 cycles per warp %.0f, waits > 3000 cycles per warp %.1f, mean wait over waits > 200: %.0f cycles" % (a[:, :, 6].mean(), a[:, :, 5].mean(), a[:, :, 4].mean(), a[:, :, 0].sum() / max(a[:, :, 5].sum(), 1)))
