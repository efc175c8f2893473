"""CPU baseline runner: the reference's CPU search path timed on the host cores.

TEST / BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline and `--impl reference` legs).

When oracle/_ref holds the reference's own compiled operators (filter_pids.cpp,
decompress_residuals.cpp, segmented_lookup.cpp, segmented_maxsim.cpp -- built from /root/reference
by oracle/build_ref.py, shipped as .so) those do the per-token work with all host threads
(`kind: "reference"`); the Python around them restates IndexScorer.rank's CPU branch
(CB/search/index_storage.py:67-184, candidate_generation.py:12-64, searcher.py:95-136) with the same
torch CPU calls, because /root/reference's Python cannot travel to the GPU box.  Without oracle/_ref
the single-threaded C restatement is used instead (`kind: "port"`).
"""
from __future__ import annotations

import os
import time

import torch

from . import build_ref
from . import plaid_oracle as po


def load_reference_ops():
    try:
        return {n: build_ref.load(n) for n in build_ref.SOURCES}
    except Exception:
        return None


class CpuSearcher:
    def __init__(self, index: po.OracleIndex, threads: int | None = None):
        self.ix = index
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.ref = load_reference_ops()
        self.kind = "reference" if self.ref is not None else "port"
        self.cores = self.threads if self.ref is not None else 1
        self.stage_s = {"candidates": 0.0, "filter": 0.0, "decompress": 0.0, "maxsim": 0.0, "other": 0.0}

    def rank(self, Q: torch.Tensor, ncells: int, thr: float, ndocs: int, query_maxlen: int = 32):
        ix, st = self.ix, self.stage_s
        t0 = time.perf_counter()
        with torch.inference_mode():
            Qc = Q[:query_maxlen]
            S = (ix.centroids @ Qc.T).contiguous()                                   # candidate_generation.py:13
            cells = torch.unique(S.topk(ncells, dim=0, sorted=False).indices.permute(1, 0).flatten())
            if self.ref is not None:
                lengths, offsets = ix.ivf_lengths[cells], ix.ivf_offsets[cells]
                pids = self.ref["segmented_lookup_cpp"].segmented_lookup_cpp(ix.ivf, cells, lengths, offsets)
            else:
                pids = torch.cat([ix.ivf[ix.ivf_offsets[c]:ix.ivf_offsets[c + 1]] for c in cells.tolist()])
            pids = torch.unique_consecutive(pids.sort().values).to(torch.int32)        # :57-60
            idx = S.max(-1).values >= thr                                             # index_storage.py:115
            t1 = time.perf_counter()
            if self.ref is not None and pids.numel() >= ndocs:
                p2 = self.ref["filter_pids_cpp"].filter_pids_cpp(pids, S, ix.codes, ix.doclens, ix.offsets, idx, ndocs)
            else:   # fewer candidates than ndocs: the C++ pops an empty heap (UB) -> restatement
                p2 = po.filter_pids(ix, pids, S, idx, ndocs)
            t2 = time.perf_counter()
            if self.ref is not None:
                D = self.ref["decompress_residuals_cpp"].decompress_residuals_cpp(
                    p2, ix.doclens, ix.offsets, ix.bucket_weights, ix.reversed_bit_map, ix.lookup, ix.residuals,
                    ix.codes, ix.centroids, ix.dim, ix.nbits)
            else:
                D = po.decompress_residuals(ix, p2)
            D = torch.nn.functional.normalize(D.to(torch.float32), p=2, dim=-1)      # index_storage.py:175
            t3 = time.perf_counter()
            lens = ix.doclens[p2.long()]
            P = D @ Q.T                                                               # colbert.py:304
            if self.ref is not None:
                scores = self.ref["segmented_maxsim_cpp"].segmented_maxsim_cpp(P, lens)
            else:
                scores = po.segmented_maxsim(P, lens)
            t4 = time.perf_counter()
            order = scores.sort(descending=True)                                      # index_storage.py:95-96
            out = (p2[order.indices].tolist(), order.values.tolist())
        t5 = time.perf_counter()
        st["candidates"] += t1 - t0
        st["filter"] += t2 - t1
        st["decompress"] += t3 - t2
        st["maxsim"] += t4 - t3
        st["other"] += t5 - t4
        return out, int(lens.sum())

    def search_all(self, Q: torch.Tensor, k: int, remove_zero_tensors: bool = False):
        """Searcher._search_all_Q: one query at a time (searcher.py:80-93).  Returns (results, T3 tokens)."""
        ncells, thr, ndocs = po.search_defaults(k)
        res, toks = [], 0
        for b in range(Q.shape[0]):
            q = po.remove_zero_rows(Q[b]) if remove_zero_tensors else Q[b]
            (pids, scores), t3 = self.rank(q, ncells, thr, ndocs)
            res.append((pids[:k], scores[:k]))
            toks += t3
        return res, toks
