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


class ShardedSearcher:
    """Wraps a per-rank `Searcher` (built with pid_range = this rank's shard)."""

    def __init__(self, searcher, group=None):
        self.searcher = searcher
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def search_batch(self, Q: torch.Tensor, k=100, remove_zero_tensors=False):
        p, s, c = self.searcher.search_batch(Q, k, remove_zero_tensors)   # global pids already
        if self.world_size == 1:
            return p, s, c
        msg = pack_lists(p, s, c)
        gathered = all_gather_lists(msg, self.world_size, self.group)
        gp, gs, gc = unpack_lists(gathered, k)
        return merge_topk(gs, gp, gc, k)
