"""pid-range sharded search over the GPUs of one box (no reference counterpart: the
reference falls back to its CPU path when world_size > 1, src/executors/FLMR_base_executor.py:883-888).

One process per GPU.  Rank g owns passages [g*N/G, (g+1)*N/G) -- its slices of codes / residuals /
doclens plus its own IVF -- and replicates the codebook and the query batch.  Every rank runs the
whole single-shard pipeline; the only exchange is one all-gather of the fixed-size per-shard top-k
lists (score f32, global pid i32, count), after which each rank merges them with the same
(score desc, pid desc) selection kernel.  Oracle: reference-per-shard + merge.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from .ops import _p, _stream


def merge_topk(scores: torch.Tensor, pids: torch.Tensor, counts: torch.Tensor, k: int):
    """scores f32 / pids i32 [G, B, k], counts i32 [G, B] (device) -> merged (pids, scores, counts) [B, k]."""
    G, B, kk = scores.shape
    assert kk == k
    dev = scores.device
    out_p = torch.empty(B, k, device=dev, dtype=torch.int32)
    out_s = torch.empty(B, k, device=dev, dtype=torch.float32)
    out_c = torch.empty(B, device=dev, dtype=torch.int32)
    ws = torch.empty(max(B * G * k, 1), device=dev, dtype=torch.int64)
    scores, pids, counts = scores.contiguous(), pids.contiguous(), counts.contiguous()   # named: outlive the launch
    _lib.call("plaid_merge_topk", _p(scores), _p(pids), _p(counts), G, B, k,
              _p(out_p), _p(out_s), _p(out_c), _p(ws), _stream())
    return out_p, out_s, out_c


def pack_lists(pids: torch.Tensor, scores: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    """One i32 message [B, 2k+1] per rank: pids | score bit patterns | count (a single all-gather)."""
    B, k = pids.shape
    msg = torch.empty(B, 2 * k + 1, device=pids.device, dtype=torch.int32)
    msg[:, :k] = pids
    msg[:, k:2 * k] = scores.view(torch.int32)
    msg[:, 2 * k] = counts
    return msg


def unpack_lists(gathered: torch.Tensor, k: int):
    """[G, B, 2k+1] -> (pids [G,B,k], scores [G,B,k], counts [G,B])."""
    pids = gathered[:, :, :k].contiguous()
    scores = gathered[:, :, k:2 * k].contiguous().view(torch.float32)
    counts = gathered[:, :, 2 * k].contiguous()
    return pids, scores, counts


def all_gather_lists(msg: torch.Tensor, world_size: int, group=None) -> torch.Tensor:
    """The path's only collective: [B, 2k+1] per rank -> [G, B, 2k+1] on every rank (NCCL over NVLink on
    the GPU box; gloo in the CPU tests)."""
    flat = torch.empty((world_size * msg.shape[0], msg.shape[1]), device=msg.device, dtype=msg.dtype)
    dist.all_gather_into_tensor(flat, msg.contiguous(), group=group)
    return flat.view(world_size, msg.shape[0], msg.shape[1])


class ListExchange:
    """The exchange step with persistent buffers: every rank's final top-k kernel writes its (local pids | score bits |
    counts) straight into the send block, ONE all_gather_into_tensor moves the blocks, and the merge kernel reads the
    receive buffer in place (adding each rank's pid base).  No pack / unpack kernels, no per-step allocations."""

    def __init__(self, B: int, k: int, device, group=None):
        self.B, self.k, self.group = B, k, group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        n = 2 * B * k + B
        self.send = torch.empty(n, device=device, dtype=torch.int32)
        self.recv = torch.empty(self.world_size * n, device=device, dtype=torch.int32)
        self.out_p = torch.empty(B, k, device=device, dtype=torch.int32)
        self.out_s = torch.empty(B, k, device=device, dtype=torch.float32)
        self.out_c = torch.empty(B, device=device, dtype=torch.int32)
        self.ws = torch.empty(max(B * self.world_size * k, 1), device=device, dtype=torch.int64)
        self.pid_bases = None

    def views(self):
        """(pids [B,k] i32, scores [B,k] f32, counts [B] i32) views of the send block."""
        B, k = self.B, self.k
        return (self.send[: B * k].view(B, k), self.send[B * k: 2 * B * k].view(torch.float32).view(B, k),
                self.send[2 * B * k:])

    def set_pid_base(self, pid_base: int):
        mine = torch.tensor([pid_base], device=self.send.device, dtype=torch.int32)
        bases = torch.empty(self.world_size, device=self.send.device, dtype=torch.int32)
        dist.all_gather_into_tensor(bases, mine, group=self.group)
        self.pid_bases = bases

    def exchange_and_merge(self):
        """all-gather of the send blocks + merge; returns (pids, scores, counts) [B, k] (global pids)."""
        dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        _lib.call("plaid_merge_topk_msg", _p(self.recv), self.world_size, self.B, self.k, _p(self.pid_bases),
                  _p(self.out_p), _p(self.out_s), _p(self.out_c), _p(self.ws), _stream())
        return self.out_p, self.out_s, self.out_c


class StageExchange:
    """Exact-global truncation of the two filter stages across pid-range shards (oracle (B)).

    The reference keeps the ndocs best stage-1 passages and the ndocs/4 best stage-2 passages of the WHOLE collection
    (filter_pids.cpp:108-123,148-157).  A shard's own top list is a superset of its share of the global one, so after
    each stage: every rank's list block ([pids | score bits | counts], written in place by the selection kernel) is
    all-gathered and every rank finds how long a prefix of its own (sorted) list belongs to the collection's top list
    (plaid_prefix_share: bisection against the other shards' lists; no merge, no copy).  The sharded search then returns exactly what one index holding the whole
    collection would -- and stage 2 / decompression / MaxSim work on 1/G of the passages per rank."""

    def __init__(self, pid_base: int, num_passages: int, device, group=None):
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        mine = torch.tensor([pid_base], device=device, dtype=torch.int32)
        self.pid_bases = torch.empty(self.world_size, device=device, dtype=torch.int32)
        dist.all_gather_into_tensor(self.pid_bases, mine, group=group)
        self._recv = {}

    def globalize(self, msg, pids, counts, Bc: int, rows: int, keep: int):
        """msg: this shard's list block of a stage ([pids | score bits | counts], lists sorted by (score, pid) descending).
        On return counts[:rows] (a view of msg) is the length of the shard's share of the collection's top-`keep` list --
        a prefix of its own list, so `pids` stays untouched."""
        del pids
        recv = self._recv.get(msg.numel())
        if recv is None:
            recv = self._recv[msg.numel()] = torch.empty(self.world_size * msg.numel(), device=msg.device, dtype=torch.int32)
        dist.all_gather_into_tensor(recv, msg, group=self.group)
        _lib.call("plaid_prefix_share", _p(recv), self.world_size, Bc, rows, keep, keep, _p(self.pid_bases), self.rank,
                  _p(counts), _stream())


class ShardedSearcher:
    """Wraps a per-rank `Searcher` (built with pid_range = this rank's shard).  `search_batch` and `_search_all_Q`
    have the single-GPU Searcher's signatures, so `search_custom_collection(sharded_searcher, ...)` works unchanged."""

    def __init__(self, searcher, group=None, mode: str = "per_shard"):
        """mode "per_shard": every shard truncates to ndocs / ndocs/4 on its own and only the final top-k lists are
        exchanged (one all-gather; oracle (A) of SURVEY.md 8e).  mode "exact": the stage lists are exchanged too
        (StageExchange), and the result equals the search of one index holding the whole collection."""
        assert mode in ("per_shard", "exact")
        self.searcher = searcher
        self.group = group
        self.mode = mode
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.config = searcher.config
        self._xchg = None
        self._qsend = self._qfull = None
        ix = searcher.ranker.index
        searcher.ranker.engine.exchange = (StageExchange(ix.pid_base, ix.num_passages, ix.device, group)
                                           if mode == "exact" and self.world_size > 1 else None)

    def _exchange(self, B, k):
        if self._xchg is None or (self._xchg.B, self._xchg.k) != (B, k):
            ix = self.searcher.ranker.index
            self._xchg = ListExchange(B, k, ix.device, self.group)
            self._xchg.set_pid_base(ix.pid_base)
        return self._xchg

    def _replicate_host_queries(self, Q: torch.Tensor) -> torch.Tensor:
        """Host query embeddings -> the whole batch on this rank's device.  Every rank is handed the same batch; instead of
        G copies of all of it over G PCIe links (33.5 MB per rank for 1024 FLMR queries, with nothing to overlap the first
        chunk's share, and the links of a box share their host side), each rank copies its 1/G slice and ONE all-gather over
        NVLink replicates it.  The search then runs on device-resident embeddings (one chunk)."""
        B, Lq, dim = Q.shape
        G, dev = self.world_size, self.searcher.ranker.index.device
        per = -(-B // G)
        if self._qfull is None or self._qfull.shape != (G * per, Lq, dim):
            self._qsend = torch.empty(per, Lq, dim, device=dev, dtype=torch.float32)
            self._qfull = torch.empty(G * per, Lq, dim, device=dev, dtype=torch.float32)
        b0 = min(B, self.rank * per)
        b1 = min(B, b0 + per)
        if b1 > b0:
            mine = Q[b0:b1].to(torch.float32).contiguous()
            if dev.type == "cuda" and not mine.is_pinned():
                mine = mine.pin_memory()
            self._qsend[: b1 - b0].copy_(mine, non_blocking=True)
        dist.all_gather_into_tensor(self._qfull, self._qsend, group=self.group)
        return self._qfull[:B]

    def search_batch(self, Q: torch.Tensor, k=100, remove_zero_tensors=False):
        """-> merged (pids, scores, counts) [B, k]; the returned tensors are reused by the next call."""
        if self.world_size == 1:
            return self.searcher.search_batch(Q, k, remove_zero_tensors)
        s = self.searcher
        s._defaults(k)
        c = s.config
        if not Q.is_cuda and Q.shape[0] >= self.world_size:
            Q = self._replicate_host_queries(Q)
        x = self._exchange(Q.shape[0], k)
        s.ranker.engine.search_batch(Q, k=k, ncells=c.ncells, centroid_score_threshold=c.centroid_score_threshold,
                                     ndocs=c.ndocs, remove_zero_rows=remove_zero_tensors, global_pids=False, out=x.views())
        return x.exchange_and_merge()

    def _search_all_Q(self, queries, Q, k, filter_fn=None, progress=True, remove_zero_tensors=False, batch_size=None):
        from .infra import Queries, Ranking
        if filter_fn is not None:
            raise NotImplementedError("filter_fn acts on shard-local candidate lists; use the per-shard Searcher")
        queries = Queries.cast(queries)
        p, sc, c = self.search_batch(Q, k, remove_zero_tensors)
        hp, hs, hc = self.searcher._host_results(p, sc, c)
        self.searcher.ranker.engine.check_flags()
        provenance = {"source": "ShardedSearcher::search_all", "queries": queries.provenance(),
                      "config": self.config.export(), "k": k, "shards": self.world_size}
        return Ranking.from_arrays(queries.keys(), hp, hs, hc, provenance)
