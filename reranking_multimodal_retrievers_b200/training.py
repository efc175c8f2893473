"""Training-time late-interaction scoring: the pieces of `FLMRModelForRetrieval.forward` that sit on
the MaxSim kernels -- `score` (colbert_score under autograd), `compute_ib_loss_new` (every query against every passage
of the batch) and `gather_tensors_from_other_gpus` (src/models/flmr/models/flmr/modeling_flmr.py:913-947,1089-1194).

Forward = the tcgen05 padded-MaxSim kernel (bf16 operands, fp32 accumulate); backward = plaid_colbert_score_backward
(arg-max recomputed on the same operands, gather for dQ, shared-memory accumulation for dD).  The loss on the small
[B, B*n_docs] score matrix and the collectives are torch / torch.distributed plumbing, as in the reference.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib, modeling, ops
from .ops import _cu, _p, _stream


class _PaddedMaxSim(torch.autograd.Function):
    """scores f32 [pairs] of Q [nQ, Lq, 128] x D_padded [n, Ld, 128]; all_pairs: [nQ * n] query-major."""

    @staticmethod
    def forward(ctx, Q, D_padded, D_mask, docs_per_query, all_pairs):
        Qf = _cu(Q.detach(), torch.float32)
        nQ, Lq, _ = Qf.shape
        n, Ld, _ = D_padded.shape
        Qb, qlens = ops.prepare_queries(Qf, remove_zero_rows=False)
        Db = modeling._as_bf16(D_padded.detach()).contiguous()
        mask = _cu(D_mask).reshape(n, Ld).ne(0).to(torch.uint8).contiguous()
        dev = Db.device
        wd = modeling._watchdog(dev)
        if all_pairs:
            scores = torch.empty(nQ, max(n, 1), device=dev, dtype=torch.float32)
            for qi in range(nQ):        # one pass over the (L2-resident) passages per query
                row = scores[qi]
                _lib.call("plaid_colbert_score_padded", _row_ptr(Qb, qi), _row_ptr(qlens, qi), 1, Qb.shape[0] - qi, Qb.shape[1],
                          _p(Db), _p(mask), n, Ld, max(n, 1), _p(row), None, Lq, _p(wd), _stream())
            scores = scores[:, :n].reshape(-1)
        else:
            scores = torch.empty(max(n, 1), device=dev, dtype=torch.float32)
            _lib.call("plaid_colbert_score_padded", _p(Qb), _p(qlens), nQ, Qb.shape[0], Qb.shape[1], _p(Db), _p(mask), n, Ld,
                      int(docs_per_query), _p(scores), None, Lq, _p(wd), _stream())
            scores = scores[:n]
        ctx.save_for_backward(Qb, qlens, Db, mask)
        ctx.meta = (nQ, Lq, n, Ld, int(docs_per_query), bool(all_pairs), Q.dtype, D_padded.dtype)
        return scores

    @staticmethod
    def backward(ctx, grad):
        Qb, qlens, Db, mask = ctx.saved_tensors
        nQ, Lq, n, Ld, dpq, all_pairs, q_dtype, d_dtype = ctx.meta
        Lq_pad = Qb.shape[1]
        dev = Db.device
        g = _cu(grad, torch.float32)
        pairs = nQ * n if all_pairs else n
        idx = torch.empty(max(pairs, 1) * Lq_pad, device=dev, dtype=torch.int32)
        dQ = torch.empty(nQ, Lq_pad, ops.DIM, device=dev, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        dD = torch.empty(n, Ld, ops.DIM, device=dev, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        if n:
            _lib.call("plaid_colbert_score_backward", _p(Qb), _p(qlens), nQ, Lq_pad, _p(Db), _p(mask), n, Ld, dpq, int(all_pairs),
                      _p(g), _p(idx), _p(dQ), _p(dD), _stream())
        return (None if dQ is None else dQ[:, :Lq].to(q_dtype), None if dD is None else dD.to(d_dtype), None, None, None)


def _row_ptr(t: torch.Tensor, row: int):
    """Raw pointer of row `row` of a contiguous tensor that the caller keeps alive."""
    return ctypes.c_void_p(t.data_ptr() + row * t.stride(0) * t.element_size())


def colbert_score(Q, D_padded, D_mask, use_gpu=True, docs_per_query=None):
    """Differentiable `colbert_score` (flmr_utils.py:33-48; scores only): Q [1 | n, Lq, 128] against the aligned
    passages, or with docs_per_query one query block per group of consecutive passages."""
    nQ, n = Q.shape[0], D_padded.shape[0]
    if docs_per_query is None:
        if nQ == 1:
            docs_per_query = max(n, 1)
        elif nQ == n:
            docs_per_query = 1
        else:
            raise ValueError(f"Q.size(0)={nQ} must be 1 or D_padded.size(0)={n}")
    return _PaddedMaxSim.apply(Q, D_padded, D_mask, int(docs_per_query), False)


def in_batch_scores(Q, D, D_mask):
    """[B, B*n_docs] MaxSim of every query against every passage of the batch (modeling_flmr.py:1098-1107), differentiable."""
    return _PaddedMaxSim.apply(Q, D, D_mask, 0, True).reshape(Q.shape[0], D.shape[0])


def compute_ib_loss_new(Q, D, D_mask, loss_fn=None):
    """In-batch-negative contrastive loss of FLMRModelForRetrieval (modeling_flmr.py:1089-1125): the positive of query i
    is passage i * (n_docs) of the batch."""
    scores = in_batch_scores(Q, D, D_mask)
    step = D.shape[0] // Q.shape[0]
    labels = torch.arange(Q.shape[0], device=scores.device) * step
    loss_fn = loss_fn or torch.nn.functional.cross_entropy
    return loss_fn(scores, labels)


def gather_tensors_from_other_gpus(query_embeddings, item_embeddings, item_mask, group=None):
    """Queries, passages and masks of all ranks concatenated in rank order; only this rank's own slices carry gradient
    (modeling_flmr.py:1127-1194: three all_gathers of detached tensors, the local block swapped back in)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return query_embeddings, item_embeddings, item_mask
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    out = []
    for t in (query_embeddings, item_embeddings, item_mask):
        local = t.detach().contiguous()
        flat = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(flat, local, group=group)
        parts = list(flat.split(local.shape[0]))
        parts[rank] = t                              # the local block keeps its autograd history
        out.append(torch.cat(parts))
    return tuple(out)
