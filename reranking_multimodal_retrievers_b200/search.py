"""Drop-in ``Searcher`` / ``IndexScorer`` for the FLMR <-> ColBERT search path.

Mirrors CB/searcher.py:25-136 and CB/search/index_storage.py:21-184: same constructor arguments,
`dense_search`, `_search_all_Q`, `rank`, `retrieve`, `score_pids`, the same k-dependent defaults and
zero-row removal -- but the index lives in HBM and a whole batch of queries goes through the CUDA
pipeline at once (`Searcher.search_batch`, used by `_search_all_Q` instead of the per-query Python
loop).  There is no CPU branch: `total_visible_gpus == 0` / no CUDA device raises.
"""
from __future__ import annotations

import os

import torch

from . import _lib, modeling, ops
from .engine import SearchEngine, search_defaults
from .index import DeviceIndex, HostIndex, load_reference_index
from .infra import ColBERTConfig, Queries, Ranking, Run


class IndexScorer:
    """GPU-resident index + scoring (CB/search/index_storage.py:21-184)."""

    # the operators the reference binds as class attributes (index_storage.py:45,58)
    filter_pids = staticmethod(ops.filter_pids)
    decompress_residuals = staticmethod(ops.decompress_residuals)

    def __init__(self, index_path, use_gpu=True, device=None, pid_range=None):
        if not use_gpu or not torch.cuda.is_available():
            raise RuntimeError("IndexScorer: this implementation has no CPU branch; a CUDA device (B200) is required")
        self.use_gpu = True
        if isinstance(index_path, (str, os.PathLike)):
            self.index_path = str(index_path)
            dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
            host = load_reference_index(self.index_path, pid_range, device=dev)    # chunk files stream straight into HBM
        else:  # an in-memory HostIndex / SyntheticIndex, or an index that is already resident (DeviceIndex)
            self.index_path = None
            host = index_path
        self.index = host if isinstance(host, DeviceIndex) else DeviceIndex(host, device)
        qml = int((getattr(host, "config", None) or {}).get("query_maxlen", ops.NQ_MAX) or ops.NQ_MAX)
        self.engine = SearchEngine(self.index, query_maxlen=qml)            # batched path (Searcher.search_batch)
        # single-query API of the reference: fp32 centroid_scores, and stage 1 by the code scan -- `score_pids` takes ANY
        # pid list / table (filter_fn, foreign callers), while the inverted-file route of the batched engine is only
        # valid for exactly the candidate bitmap the preceding `stage_candidates` left in the workspace
        self.engine1 = SearchEngine(self.index, s_dtype=torch.float32, ivf_stage1=False, query_maxlen=qml)
        self.doclens = self.index.doclens
        self.num_embeddings = self.index.num_embeddings
        self.num_partitions = self.index.num_centroids
        self._emb2pid = None

    def set_query_maxlen(self, query_maxlen):
        """`config.query_maxlen` of the searcher (CB/search/index_storage.py:77: `Q[:, :config.query_maxlen]`)."""
        if not 1 <= int(query_maxlen) <= ops.NQ_MAX:
            raise _lib.PlaidError(f"query_maxlen={query_maxlen}: the candidate stage supports 1..{ops.NQ_MAX} query tokens "
                                  "(one per lane); longer candidate-stage queries are not silently truncated")
        for eng in (self.engine, self.engine1):
            eng.query_maxlen = int(query_maxlen)

    # ---- single-query API of the reference -------------------------------------------------
    def _prep(self, config, Q, k_hint=None):
        Q = Q if Q.dim() == 3 else Q.unsqueeze(0)
        assert Q.size(0) == 1, "IndexScorer.rank/retrieve take one query [1, Lq, dim] (searcher.py:81)"
        qml = getattr(config, "query_maxlen", None)
        if qml:
            self.set_query_maxlen(qml)
        ncells, thr, ndocs = config.ncells, config.centroid_score_threshold, config.ndocs
        d = search_defaults(k_hint or 10)
        ncells = d[0] if ncells is None else ncells
        thr = d[1] if thr is None else thr
        ndocs = d[2] if ndocs is None else ndocs
        Lq_pad = ((Q.shape[1] + 31) // 32) * 32
        ws = self.engine1._workspace(4, Lq_pad, int(ncells), int(ndocs), int(ndocs) // 4)
        return Q, ws, int(ncells), float(thr), int(ndocs), Lq_pad

    def retrieve(self, config, Q):
        """(candidate pids i32 sorted unique, centroid_scores f32 [C, nq]) -- index_storage.py:67-80.
        Independent tensors, like the reference's: a later retrieve/rank does not change them."""
        Q, ws, ncells, thr, ndocs, Lq_pad = self._prep(config, Q)
        Qd = ops._cu(Q, torch.float32)
        with torch.cuda.device(self.index.device):
            self.engine1.stage_candidates(ws, Qd, Lq_pad, ncells, thr, False, 4)
        n = int(ws["cand_counts"][0].item())
        nq = min(int(ws["qlens"][0].item()), self.engine1.query_maxlen)
        return ws["cand_pids"][0, :n].clone(), ws["S"][0, :, :nq].clone()

    def score_pids(self, config, Q, pids, centroid_scores, batch_size=None):
        """(scores f32, pids i32) of the ndocs/4 passages surviving the two-stage filter
        (index_storage.py:100-184).  `pids` is ANY list of local pids (what `retrieve` returned, a `filter_fn`'s
        subset of it, or the caller's own) and `centroid_scores` any f32 [C, nq] table: both are installed into
        the workspace and the pruning mask is rebuilt from the table (index_storage.py:115)."""
        Q, ws, ncells, thr, ndocs, Lq_pad = self._prep(config, Q)
        S_ws = ws["S"][0]
        nq = centroid_scores.shape[1]
        if nq > ops.NQ_MAX:
            raise _lib.PlaidError(f"score_pids: centroid_scores has {nq} query-token columns > {ops.NQ_MAX}")
        with torch.cuda.device(self.index.device):
            Qb, qlens, Qh = ops.prepare_queries(ops._cu(Q, torch.float32), False, Lq_pad, 4, with_f16=True)
            ws["Qb"].copy_(Qb)
            ws["Qh"].copy_(Qh)
            ws["qlens"].copy_(qlens)
            S_ws.zero_()
            S_ws[:, :nq] = ops._cu(centroid_scores, torch.float32)
            idx = S_ws[:, :nq].max(-1).values >= thr
            ws["idx_bits"][0] = ops.pack_idx_bits(idx)
            pids = ops._cu(pids, torch.int32).reshape(-1)
            n = pids.numel()
            if n > ws["cand_stride"]:
                raise ValueError(f"{n} candidate pids exceed the workspace ({ws['cand_stride']})")
            if n and (int(pids.min()) < 0 or int(pids.max()) >= self.index.num_passages):
                raise _lib.PlaidError("score_pids: pid outside this index shard")
            ws["cand_pids"][0, :n] = pids
            ws["cand_counts"][0] = n
            # the candidate stage of the filter sees min(qlen, nq, query_maxlen) tokens, as the reference's table has nq columns
            saved = self.engine1.query_maxlen
            self.engine1.query_maxlen = min(saved, nq)
            try:
                self.engine1.stage_rank(ws, 1, Lq_pad, ndocs, ndocs // 4, 4)
            finally:
                self.engine1.query_maxlen = saved
        m = int(ws["s2_counts"][0].item())
        return ws["scores"][0, :m].clone(), ws["s2_pids"][0, :m].clone()

    def rank(self, config, Q, filter_fn=None, batch_size=None):
        """(pids list, scores list), best first (index_storage.py:86-98)."""
        with torch.inference_mode():
            pids, centroid_scores = self.retrieve(config, Q)
            if filter_fn is not None:
                pids = filter_fn(pids)
            scores, pids = self.score_pids(config, Q, pids, centroid_scores, batch_size=batch_size)
            op, os_ = ops.select_top(pids, scores, max(int(pids.numel()), 1))
            op = op + self.index.pid_base
            return op.tolist(), os_.tolist()

    def lookup_eids(self, embedding_ids, codes=None, out_device="cuda"):
        """fp16 normalised embeddings [n, dim] of the listed token ids (index_storage.py:61-62 ->
        residual_embeddings_strided.py:24-28 -> ResidualCodec.decompress, GPU branch)."""
        ix = self.index
        eids = ops._cu(embedding_ids, torch.int64).reshape(-1)
        if eids.numel() and (int(eids.min()) < 0 or int(eids.max()) >= ix.num_embeddings):
            raise _lib.PlaidError("lookup_eids: embedding id outside this index shard")
        codes = ix.codes[eids] if codes is None else ops._cu(codes, torch.int32)
        residuals = ix.residuals[eids]
        return ops.codec_decompress_residuals(residuals, ix.bucket_weights.half(), ix.reversed_bit_map, ix.lookup_table,
                                              codes, ix.centroids_f16, ix.dim, ix.nbits, normalize=True)

    def embedding_ids_to_pids(self, embedding_ids):
        """Unique passages owning the listed token ids (index_storage.py:82-84; LOCAL pids of this shard)."""
        if self._emb2pid is None:
            ix = self.index
            self._emb2pid = torch.repeat_interleave(torch.arange(ix.num_passages, device=ix.device, dtype=torch.int32),
                                                    ix.doclens)
        return torch.unique(self._emb2pid[ops._cu(embedding_ids, torch.int64).reshape(-1)], sorted=False)

    def lookup_pids(self, passage_ids, out_device="cuda", return_mask=False):
        """(D_packed f32 normalised [sum len, dim], lengths) for local pids (index_storage.py:64-65)."""
        ix = self.index
        pids = ops._cu(passage_ids, torch.int32)
        D = ops.decompress_residuals(pids, ix.doclens, ix.offsets, ix.bucket_weights, ix.reversed_bit_map,
                                     ix.lookup_table, ix.residuals, ix.codes, ix.centroids_f32, ix.dim, ix.nbits)
        return torch.nn.functional.normalize(D, p=2, dim=-1), ix.doclens[pids.long()]


class Searcher:
    """CB/searcher.py:24-136 for precomputed query embeddings (the only mode this fork supports:
    `Searcher.encode/search/search_all` need a text checkpoint the fork never loads)."""

    def __init__(self, index, checkpoint=None, collection=None, config=None, disable_gpu=True, device=None,
                 pid_range=None):
        initial = ColBERTConfig.from_existing(config, Run().config)
        if config is not None:
            initial.total_visible_gpus = config.total_visible_gpus
        if isinstance(index, (str, os.PathLike)):
            self.index = index if os.path.isdir(index) else os.path.join(initial.index_root_, index)
            self.index_config = ColBERTConfig.load_from_index(self.index)
        else:
            self.index = index
            self.index_config = ColBERTConfig(**(getattr(index, "config", None) or {}))
        self.checkpoint = checkpoint or self.index_config.checkpoint
        self.config = ColBERTConfig.from_existing(self.index_config, initial)
        # search knobs stored in the index metadata are indexing-time leftovers: the reference's
        # defaults are None ("choose from k"), keep that unless the caller set them
        for knob in ("ncells", "centroid_score_threshold", "ndocs"):
            if config is None or knob not in config._assigned:
                self.config.configure(**{knob: None})
        self.collection = collection or self.config.collection
        if self.config.total_visible_gpus == 0:
            raise RuntimeError("Searcher: total_visible_gpus=0 selects the reference's CPU branch, which this "
                               "B200 implementation does not have")
        self.ranker = IndexScorer(self.index, True, device=device, pid_range=pid_range)

    def configure(self, **kw):
        self.config.configure(**kw)

    # ---- text entry points of the reference: unsupported in this fork as well --------------
    def encode(self, text):
        raise NotImplementedError("text encoding needs a checkpoint this fork never loads (searcher.py:42-66); "
                                  "pass query embeddings to dense_search/_search_all_Q")

    def search(self, text, k=10, filter_fn=None):
        return self.dense_search(self.encode(text), k, filter_fn=filter_fn)

    def search_all(self, queries, k=10, filter_fn=None):
        queries = Queries.cast(queries)
        return self._search_all_Q(queries, self.encode(list(queries.values())), k, filter_fn=filter_fn)

    # ---- the path FLMR uses -----------------------------------------------------------------
    def _defaults(self, k):
        """k-dependent defaults, assigned once like searcher.py:96-122."""
        ncells, thr, ndocs = search_defaults(k)
        if self.config.ncells is None:
            self.configure(ncells=ncells)
        if self.config.centroid_score_threshold is None:
            self.configure(centroid_score_threshold=thr)
        if self.config.ndocs is None:
            self.configure(ndocs=ndocs)
        # `Q[:, :config.query_maxlen]` drives candidate generation (index_storage.py:77); > 32 raises
        self.ranker.set_query_maxlen(self.config.query_maxlen or ops.NQ_MAX)

    def search_batch(self, Q: torch.Tensor, k=10, remove_zero_tensors=False):
        """Q f32 [B, Lq, dim] -> (pids i32 [B, k], scores f32 [B, k], counts i32 [B]) device tensors."""
        self._defaults(k)
        c = self.config
        return self.ranker.engine.search_batch(Q, k=k, ncells=c.ncells,
                                               centroid_score_threshold=c.centroid_score_threshold, ndocs=c.ndocs,
                                               remove_zero_rows=remove_zero_tensors)

    def dense_search(self, Q: torch.Tensor, k=10, filter_fn=None, remove_zero_tensors=False, batch_size=None):
        """One query [1, Lq, dim] -> (pids, ranks, scores) python lists (searcher.py:95-136)."""
        self._defaults(k)
        if filter_fn is not None:
            if remove_zero_tensors:
                Q = Q[torch.abs(Q).sum(dim=-1) > 0].unsqueeze(0)
            pids, scores = self.ranker.rank(self.config, Q, filter_fn=filter_fn, batch_size=batch_size)
            return pids[:k], list(range(1, k + 1)), scores[:k]
        p, s, c = self.search_batch(Q if Q.dim() == 3 else Q.unsqueeze(0), k, remove_zero_tensors)
        n = int(c[0].item())
        return p[0, :n].tolist(), list(range(1, k + 1)), s[0, :n].tolist()

    def _search_all_Q(self, queries, Q, k, filter_fn=None, progress=True, remove_zero_tensors=False, batch_size=None):
        queries = Queries.cast(queries)
        if filter_fn is not None:
            rows = [list(zip(*self.dense_search(Q[i:i + 1], k, filter_fn=filter_fn,
                                                remove_zero_tensors=remove_zero_tensors, batch_size=batch_size)))
                    for i in range(Q.size(0))]
        else:
            p, s, c = self.search_batch(Q, k, remove_zero_tensors)
            hp, hs, hc = self._host_results(p, s, c)              # one pinned D2H + one synchronisation
            self.ranker.engine.check_flags()
            provenance = {"source": "Searcher::search_all", "queries": queries.provenance(),
                          "config": self.config.export(), "k": k}
            return Ranking.from_arrays(queries.keys(), hp, hs, hc, provenance)
        data = {qid: val for qid, val in zip(queries.keys(), rows)}
        provenance = {"source": "Searcher::search_all", "queries": queries.provenance(),
                      "config": self.config.export(), "k": k}
        return Ranking(data=data, provenance=provenance)

    def _host_results(self, p, s, c):
        """Device lists -> numpy arrays through pinned staging buffers (fresh numpy copies: the Ranking owns them)."""
        B, k = p.shape
        key = (B, k)
        if getattr(self, "_pinned_key", None) != key:
            self._pinned = (torch.empty(B, k, dtype=torch.int32).pin_memory(), torch.empty(B, k, dtype=torch.float32).pin_memory(),
                            torch.empty(B, dtype=torch.int32).pin_memory())
            self._pinned_key = key
        hp, hs, hc = self._pinned
        hp.copy_(p, non_blocking=True)
        hs.copy_(s, non_blocking=True)
        hc.copy_(c, non_blocking=True)
        torch.cuda.current_stream(p.device).synchronize()
        return hp.numpy().copy(), hs.numpy().copy(), hc.numpy().copy()


# re-exported so `from ... import colbert_score` works like `colbert.modeling.colbert`
colbert_score = modeling.colbert_score
colbert_score_packed = modeling.colbert_score_packed
colbert_score_reduce = modeling.colbert_score_reduce
