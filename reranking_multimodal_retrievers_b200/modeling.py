"""Late-interaction score functions with the reference's names and signatures.

  colbert_score_reduce / colbert_score / colbert_score_packed   CB/modeling/colbert.py:235-311
  flmr_colbert_score_reduce / flmr_colbert_score                src/models/flmr/models/flmr/flmr_utils.py:22-48
                                                                 (same maths, return (scores, scores_padded))

Operands are rounded to bf16 and accumulated in fp32 on the tensor cores (tcgen05); the per-passage
max and the sum over query tokens happen in the kernel epilogue.  Scores come back in fp32 (the
reference returns them in D's dtype because it casts Q to it, colbert.py:284).  Only the 'colbert'
interaction is implemented (the 'flipr' branch, colbert.py:248-261, is unused by this repo's configs).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, ops
from .ops import _cu, _p, _stream

_FLAGS = {}


def _watchdog(dev):
    key = (dev.type, dev.index)
    if key not in _FLAGS:
        _FLAGS[key] = torch.zeros(1, device=dev, dtype=torch.int32)
    return _FLAGS[key]


def _check_interaction(config):
    inter = getattr(config, "interaction", "colbert") if config is not None else "colbert"
    if inter != "colbert":
        raise NotImplementedError(f"interaction={inter!r}: only 'colbert' is implemented on this path")


def _as_bf16(x: torch.Tensor) -> torch.Tensor:
    x = _cu(x)
    if x.dtype == torch.bfloat16:
        return x
    if x.dtype == torch.float32:
        return ops.to_bf16(x)
    return x.to(torch.bfloat16)


def colbert_score_reduce(scores_padded, D_mask, config=None):
    """-9999 fill of padded positions, max over passage tokens, sum over query tokens (colbert.py:237-263)."""
    _check_interaction(config)
    sp = _cu(scores_padded, torch.float32)
    n, Ld, Lq = sp.shape
    mask = _cu(D_mask).reshape(n, Ld).ne(0).to(torch.uint8).contiguous()
    out = torch.empty(max(n, 1), device=sp.device, dtype=torch.float32)
    _lib.call("plaid_colbert_score_reduce", _p(sp), _p(mask), n, Ld, Lq, _p(out), _stream())
    return out[:n]


_HOST_Q_CHUNK_BYTES = 16 << 20     # host query batches of colbert_score cross PCIe in pieces of about this size
_COPY_STREAMS = {}


def _padded_scores_host_queries(Q, D_padded, D_mask, docs_per_query):
    """colbert_score with a large batch of HOST query matrices against device-resident passages (the cross-encoder
    hand-off shape: 4096 queries x 100 passages): the queries cross PCIe in ~16 MB pieces on a copy stream, one piece ahead
    of the MaxSim of the previous one, instead of as one exposed 134 MB copy."""
    nQ, Lq, dim = Q.shape
    n, Ld, _ = D_padded.shape
    dev = D_padded.device
    Qh = Q.to(torch.float32).contiguous()
    if not Qh.is_pinned():
        Qh = Qh.pin_memory()
    chunk = max(4, (_HOST_Q_CHUNK_BYTES // (Lq * dim * 4) // 4) * 4)
    Db = _as_bf16(D_padded).contiguous()
    mask = _cu(D_mask).reshape(n, Ld).ne(0).to(torch.uint8).contiguous()
    scores = torch.empty(max(n, 1), device=dev, dtype=torch.float32)
    wd = _watchdog(dev)
    main = torch.cuda.current_stream(dev)
    key = (dev.type, dev.index)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=dev)
    copy = _COPY_STREAMS[key]
    bufs = [torch.empty(chunk, Lq, dim, device=dev, dtype=torch.float32) for _ in range(2)]
    copy.wait_stream(main)                                   # the staging buffers were allocated on the compute stream
    freed = [None, None]
    ready = [None, None]

    def issue(ci):
        q0 = ci * chunk
        if q0 >= nQ or q0 * docs_per_query >= n:
            return
        q1 = min(nQ, q0 + chunk)
        slot = ci & 1
        if freed[slot] is not None:
            copy.wait_event(freed[slot])
        with torch.cuda.stream(copy):
            bufs[slot][: q1 - q0].copy_(Qh[q0:q1], non_blocking=True)
            ready[slot] = torch.cuda.Event()
            ready[slot].record(copy)

    issue(0)
    for ci in range((nQ + chunk - 1) // chunk):
        q0 = ci * chunk
        d0 = q0 * docs_per_query
        if d0 >= n:
            break
        q1 = min(nQ, q0 + chunk)
        d1 = min(n, q1 * docs_per_query)
        slot = ci & 1
        issue(ci + 1)
        main.wait_event(ready[slot])
        Qb, qlens = ops.prepare_queries(bufs[slot][: q1 - q0], remove_zero_rows=False)
        freed[slot] = torch.cuda.Event()
        freed[slot].record(main)
        _lib.call("plaid_colbert_score_padded", _p(Qb), _p(qlens), q1 - q0, Qb.shape[0], Qb.shape[1], _p(Db[d0:d1]),
                  _p(mask[d0:d1]), d1 - d0, Ld, int(docs_per_query), _p(scores[d0:d1]), None, Lq, _p(wd), _stream())
    return scores[:n]


def _padded_scores(Q, D_padded, D_mask, return_raw, docs_per_query=None):
    if (not return_raw and docs_per_query is not None and torch.is_tensor(Q) and not Q.is_cuda and Q.dim() == 3
            and torch.is_tensor(D_padded) and D_padded.is_cuda and D_padded.dim() == 3
            and Q.shape[0] * Q.shape[1] * Q.shape[2] * 4 >= 2 * _HOST_Q_CHUNK_BYTES
            and Q.shape[0] * int(docs_per_query) >= D_padded.shape[0]):
        return _padded_scores_host_queries(Q, D_padded, D_mask, int(docs_per_query)), None
    Q = _cu(Q)
    if Q.dim() != 3 or D_padded.dim() != 3:
        raise ValueError("colbert_score expects Q [1|n, Lq, dim] and D_padded [n, Ld, dim]")
    n, Ld, dim = D_padded.shape
    nQ, Lq, _ = Q.shape
    if docs_per_query is None:
        if nQ == 1:
            docs_per_query = max(n, 1)
        elif nQ == n:
            docs_per_query = 1  # "each query matrix is compared against the aligned passage" (colbert.py:276-279)
        else:
            raise ValueError(f"Q.size(0)={nQ} must be 1 or D_padded.size(0)={n} (colbert.py:281)")
    Qb, qlens = ops.prepare_queries(Q.float(), remove_zero_rows=False)
    Db = _as_bf16(D_padded).contiguous()
    mask = _cu(D_mask).reshape(n, Ld).ne(0).to(torch.uint8).contiguous()
    dev = Db.device
    scores = torch.empty(max(n, 1), device=dev, dtype=torch.float32)
    raw = torch.empty(n, Ld, Lq, device=dev, dtype=torch.float32) if return_raw else None
    wd = _watchdog(dev)
    _lib.call("plaid_colbert_score_padded", _p(Qb), _p(qlens), nQ, Qb.shape[0], Qb.shape[1], _p(Db), _p(mask), n, Ld,
              int(docs_per_query), _p(scores), _p(raw), Lq, _p(wd), _stream())
    return scores[:n], raw


def colbert_score(Q, D_padded, D_mask, config=None, use_gpu=True, docs_per_query=None):
    """Padded MaxSim (colbert.py:268-286).  `docs_per_query` (extension) lets Q hold one row block per
    query while D_padded holds docs_per_query consecutive passages for each."""
    _check_interaction(config)
    return _padded_scores(Q, D_padded, D_mask, False, docs_per_query)[0]


def colbert_score_packed(Q, D_packed, D_lengths, config=None):
    """Packed MaxSim for ONE query (colbert.py:289-311 CPU branch + segmented_maxsim.cpp): per-token
    max clamped at 0, summed over query tokens."""
    _check_interaction(config)
    Q = _cu(Q)
    if Q.dim() == 2:
        Q = Q.unsqueeze(0)
    assert Q.size(0) == 1, "colbert_score_packed scores one query against a packed set of passages"
    Qb, qlens = ops.prepare_queries(Q.float(), remove_zero_rows=False)
    Db = _as_bf16(D_packed).contiguous()
    dev = Db.device
    lengths = _cu(D_lengths, torch.int64)
    nd = lengths.numel()
    T = Db.shape[0]
    tok_offsets = torch.zeros(nd + 1, device=dev, dtype=torch.int32)
    tok_offsets[1:] = torch.cumsum(lengths, 0).to(torch.int32)
    counts = torch.tensor([nd], device=dev, dtype=torch.int32)
    scores = torch.zeros(max(nd, 1), device=dev, dtype=torch.float32)
    if nd == 0 or T == 0:      # nothing to contract: every (empty) passage scores sum_k max(0, -) = 0
        return scores[:nd]
    wd = _watchdog(dev)
    _lib.call("plaid_maxsim_packed", _p(Qb), _p(qlens), 1, Qb.shape[0], Qb.shape[1], _p(Db), _p(tok_offsets), _p(counts),
              max(nd, 1), max(T, 1), 1, 0, 0, _p(scores), _p(wd), _stream())
    return scores[:nd]


# ---- FLMR copies (flmr_utils.py:22-48): same maths, return the masked similarity matrix too ----
def flmr_colbert_score_reduce(scores_padded, D_mask):
    sp = _cu(scores_padded, torch.float32).clone()
    pad = ~_cu(D_mask).reshape(sp.size(0), sp.size(1)).bool()
    scores = colbert_score_reduce(sp, D_mask)
    sp[pad] = -9999
    return scores, sp


def flmr_colbert_score(Q, D_padded, D_mask, use_gpu=True, docs_per_query=None):
    """(scores [n], scores_padded [n, Ld, Lq]) -- the second tensor is the `scores_raw` handoff the
    rerankers read (src/models/flmr/models/flmr/modeling_flmr.py:936, 1601-1602)."""
    return _padded_scores(Q, D_padded, D_mask, True, docs_per_query)
