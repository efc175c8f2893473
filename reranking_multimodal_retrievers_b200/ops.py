"""Reference-named operators over libplaid_b200.so.

These mirror, argument for argument, the native operators the reference binds as class attributes
(SURVEY.md 8b):

  IndexScorer.filter_pids            <- filter_pids_cpp           (CB/search/filter_pids.cpp:126-164)
  IndexScorer.decompress_residuals   <- decompress_residuals_cpp  (CB/search/decompress_residuals.cpp:80-155)
  ColBERT.segmented_maxsim           <- segmented_maxsim_cpp      (CB/modeling/segmented_maxsim.cpp:49-93)
  StridedTensor.segmented_lookup     <- segmented_lookup_cpp      (CB/search/segmented_lookup.cpp:51-125)
  ResidualCodec.decompress_residuals <- decompress_residuals_cpp  (CB/indexing/codecs/decompress_residuals.cu:8-75, GPU form)
  ResidualCodec.packbits             <- packbits_cpp              (CB/indexing/codecs/packbits.cu:10-57)

Inputs are torch CUDA tensors (CPU tensors are moved to the current CUDA device first, since the
reference calls these with CPU tensors); outputs are CUDA tensors.  torch is used only to own
memory and streams -- all arithmetic happens in the CUDA library; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

NQ_MAX = 32
DIM = 128


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("reranking_multimodal_retrievers_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _cu(t: torch.Tensor, dtype=None) -> torch.Tensor:
    """contiguous, 16-byte aligned tensor on the current CUDA device (optionally cast)."""
    if (t.is_cuda and (dtype is None or t.dtype == dtype) and t.is_contiguous() and not (t.data_ptr() & 15)
            and t.device.index == torch.cuda.current_device()):
        return t            # already what the kernels want: the common case of the operator calls, kept off the slow path
    t = t.to(device=_dev(), dtype=dtype if dtype is not None else t.dtype, non_blocking=True)
    t = t.contiguous()
    if t.data_ptr() % 16:   # a view into the middle of a buffer: the kernels use 128-bit loads
        t = t.clone()
    return t


def _p(t):
    """Raw device pointer.  Only ever applied to a NAMED tensor: a temporary would be returned to the
    caching allocator as soon as this call returns and could be handed to the next allocation."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def pack_idx_bits(idx: torch.Tensor) -> torch.Tensor:
    """bool [.., C] -> u32 words [.., C/32] stored as int32 (bit c%32 of word c/32)."""
    C = idx.shape[-1]
    assert C % 32 == 0, "number of centroids must be a multiple of 32"
    w = idx.reshape(*idx.shape[:-1], C // 32, 32).to(torch.int64)
    sh = torch.arange(32, device=idx.device, dtype=torch.int64)
    words = (w << sh).sum(-1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words)
    return words.to(torch.int32).contiguous()


def unpack_idx_bits(words: torch.Tensor, C: int) -> torch.Tensor:
    sh = torch.arange(32, device=words.device, dtype=torch.int64)
    bits = (words.to(torch.int64).unsqueeze(-1) >> sh) & 1
    return bits.reshape(*words.shape[:-1], -1)[..., :C].bool()


def check_codes(codes: torch.Tensor, C: int):
    """The filter kernels index the pruning bitmap and the S table with the codes unchecked (the reference
    asserts `code < ncentroids` and aborts, filter_pids.cpp:47); validate once on the host side instead."""
    if codes.numel() and (int(codes.max()) >= C or int(codes.min()) < 0):
        raise _lib.PlaidError(f"centroid codes outside [0, {C})")


def pad_centroid_scores(centroid_scores: torch.Tensor) -> torch.Tensor:
    """[C, nq] (nq <= 32) -> the kernels' [C, 32] row layout."""
    C, nq = centroid_scores.shape
    if nq > NQ_MAX:
        raise _lib.PlaidError(f"filter_pids: {nq} candidate-stage query tokens > {NQ_MAX} (query_maxlen)")
    S = centroid_scores.float()
    if nq == NQ_MAX:
        return S.contiguous()
    out = torch.zeros(C, NQ_MAX, device=S.device, dtype=torch.float32)
    out[:, :nq] = S
    return out


# --------------------------------------------------------------------------------------------- filter_pids
def approx_scores(pids, centroid_scores, codes, offsets, idx=None, table_f16=False):
    """Per-passage approximate score (filter_pids.cpp:27-72) for ONE query; idx=None = all centroids.
    table_f16: hand the kernel the table in fp16, the precision of the reference's GPU branch
    (candidate_generation.py:52) and the engine's default; the sums stay sequential fp32."""
    pids = _cu(pids, torch.int32)
    S = pad_centroid_scores(_cu(centroid_scores))
    if table_f16:
        S = S.half()
    C, nq = centroid_scores.shape
    n = pids.numel()
    counts = torch.tensor([n], device=pids.device, dtype=torch.int32)
    qlens = torch.tensor([nq], device=pids.device, dtype=torch.int32)
    bits = pack_idx_bits(_cu(idx).bool()) if idx is not None else None
    out = torch.empty(max(n, 1), device=pids.device, dtype=torch.float32)
    codes, offsets = _cu(codes, torch.int32), _cu(offsets, torch.int64)   # named: must outlive the launch
    check_codes(codes, C)
    _lib.call("plaid_approx_scores", _p(pids), _p(counts), 1, n, _p(S), int(table_f16), _p(qlens), _p(bits), C,
              _p(codes), _p(offsets), _p(out), _stream())
    return out[:n]


def select_top(pids, scores, keep):
    """(score, pid)-descending top-`keep` for ONE list; returns (pids, scores)."""
    pids = _cu(pids, torch.int32)
    scores = _cu(scores, torch.float32)
    n = pids.numel()
    dev = pids.device
    counts = torch.tensor([n], device=dev, dtype=torch.int32)
    op = torch.empty(keep, device=dev, dtype=torch.int32)
    os_ = torch.empty(keep, device=dev, dtype=torch.float32)
    oc = torch.empty(1, device=dev, dtype=torch.int32)
    ws = torch.empty(max(n, 1), device=dev, dtype=torch.int64)
    _lib.call("plaid_select_top", _p(pids), _p(scores), _p(counts), 1, n, keep, _p(op), _p(os_), _p(oc), keep,
              _p(ws), _stream())
    m = min(n, keep)
    return op[:m], os_[:m]


def filter_pids(pids, centroid_scores, codes, doclens, offsets, idx, nfiltered_docs, return_stages=False):
    """Drop-in for ``IndexScorer.filter_pids`` (index_storage.py:153-156): i32 pids of the
    min(n, ndocs/4) best passages, (score, pid) descending."""
    del doclens  # lengths are offsets[pid+1]-offsets[pid] (strided_tensor_core.py:30-31)
    pids = _cu(pids, torch.int32)
    dev = pids.device
    C, nq = centroid_scores.shape
    S = pad_centroid_scores(_cu(centroid_scores))
    bits = pack_idx_bits(_cu(idx).bool())
    n, ndocs = pids.numel(), int(nfiltered_docs)
    stride = max(n, ndocs, 1)
    counts = torch.tensor([n], device=dev, dtype=torch.int32)
    qlens = torch.tensor([nq], device=dev, dtype=torch.int32)
    pid_buf = torch.full((stride,), -1, device=dev, dtype=torch.int32)
    pid_buf[:n] = pids
    ws_scores = torch.empty(stride, device=dev, dtype=torch.float32)
    ws_keys = torch.empty(stride, device=dev, dtype=torch.int64)
    s1p = torch.empty(ndocs, device=dev, dtype=torch.int32)
    s1s = torch.empty(ndocs, device=dev, dtype=torch.float32)
    s1c = torch.empty(1, device=dev, dtype=torch.int32)
    s2p = torch.empty(ndocs // 4, device=dev, dtype=torch.int32)
    s2s = torch.empty(ndocs // 4, device=dev, dtype=torch.float32)
    s2c = torch.empty(1, device=dev, dtype=torch.int32)
    codes, offsets = _cu(codes, torch.int32), _cu(offsets, torch.int64)   # named: must outlive the launch
    check_codes(codes, C)
    _lib.call("plaid_filter_pids", _p(pid_buf), _p(counts), 1, stride, _p(S), 0, _p(qlens), _p(bits), C,
              _p(codes), _p(offsets), ndocs, _p(ws_scores), _p(ws_keys),
              _p(s1p), _p(s1s), _p(s1c), _p(s2p), _p(s2s), _p(s2c), _stream())
    n1, n2 = min(n, ndocs), min(n, ndocs // 4)
    if return_stages:
        return s2p[:n2], (s1p[:n1], s1s[:n1], s2s[:n2])
    return s2p[:n2]


# --------------------------------------------------------------------------------------------- decompression
_WT_CACHE = []      # [(tensors, versions, device, nbits, W, event, stream)], most recent first


def build_weight_table(bucket_weights, reversed_bit_map, lookup, nbits):
    """f32 [256 * 8/nbits]: packed residual byte -> the bucket weights it expands to (`bucket_weights[lookup[
    reversed_bit_map[x]]]`, residual.py:54-89).  The operators take the three tables on every call, as the reference's
    do; the derived table is kept for the same three tensor OBJECTS at unchanged versions (the cache holds them, so their
    storage cannot be recycled under the key) -- a 2^15-token operator call, the batch ResidualCodec.decompress uses
    (residual.py:246), otherwise spends more host time rebuilding this table than the device spends decoding."""
    dev = _dev()
    key = (bucket_weights, reversed_bit_map, lookup)
    try:
        vers = tuple(t._version for t in key)
    except (RuntimeError, AttributeError):     # inference tensors carry no version counter: not cacheable
        vers = None
    if vers is not None:
        for ent in _WT_CACHE:
            if ent[3] == int(nbits) and ent[2] == dev and ent[1] == vers and all(a is b for a, b in zip(ent[0], key)):
                st = torch.cuda.current_stream(dev)
                if st != ent[6]:
                    st.wait_event(ent[5])      # built on another stream
                return ent[4]
    W = torch.empty(256 * (8 // nbits), device=dev, dtype=torch.float32)
    bw, rbm, lut = _cu(bucket_weights, torch.float32), _cu(reversed_bit_map, torch.uint8), _cu(lookup, torch.uint8)
    _lib.call("plaid_build_weight_table", _p(bw), _p(rbm), _p(lut), int(nbits), _p(W), _stream())
    if vers is not None:
        ev = torch.cuda.Event()
        st = torch.cuda.current_stream(dev)
        ev.record(st)
        _WT_CACHE.insert(0, (key, vers, dev, int(nbits), W, ev, st))
        del _WT_CACHE[4:]
    return W


def decompress_residuals(pids, lengths, offsets, bucket_weights, reversed_bit_map, bucket_weight_combinations,
                         binary_residuals, codes, centroids, dim, nbits):
    """Drop-in for ``IndexScorer.decompress_residuals`` (index_storage.py:162-174): f32
    [sum(lengths[pids]), dim], bucket weight + centroid, rows packed in pid order."""
    if dim != DIM:
        raise _lib.PlaidError(f"decompress_residuals: dim={dim}, the kernels are built for dim={DIM}")
    pids = _cu(pids, torch.int32)
    dev = pids.device
    offsets = _cu(offsets, torch.int64)
    lengths = _cu(lengths, torch.int64)
    lens = lengths[pids.long()]
    out_offsets = torch.zeros(pids.numel() + 1, device=dev, dtype=torch.int64)
    out_offsets[1:] = torch.cumsum(lens, 0)
    total = int(out_offsets[-1].item()) if pids.numel() else 0
    # the kernel reads lengths as offsets[pid+1]-offsets[pid]; make that hold for a bare cumsum
    if offsets.numel() == lengths.numel():
        offsets = torch.cat((offsets, (offsets[-1:] + lengths[-1:])))
    W = build_weight_table(bucket_weights, reversed_bit_map, bucket_weight_combinations, nbits)
    cent = _cu(centroids, torch.float32)
    out = torch.empty(max(total, 1), dim, device=dev, dtype=torch.float32)
    res, codes = _cu(binary_residuals, torch.uint8), _cu(codes, torch.int32)   # named: must outlive the launch
    if pids.numel():
        _lib.call("plaid_decompress_residuals", _p(pids), pids.numel(), _p(offsets), _p(out_offsets), _p(W),
                  _p(res), _p(codes), _p(cent), cent.shape[0], int(nbits), _p(out), _stream())
    return out[:total]


def codec_decompress_residuals(binary_residuals, bucket_weights, reversed_bit_map, bucket_weight_combinations, codes,
                               centroids, dim, nbits, normalize=False):
    """Drop-in for ``ResidualCodec.decompress_residuals`` -- the GPU-branch operator (residual.py:115,250-260;
    decompress_residuals.cu:8-75): token rows without pid indirection, fp16 [n, dim] = half bucket weight + half
    centroid (one half add per element, bit-exact with the reference kernel).  normalize=True applies
    ``ResidualCodec.decompress``'s `F.normalize(..).half()` in the same launch (residual.py:272-273)."""
    if dim != DIM:
        raise _lib.PlaidError(f"decompress_residuals: dim={dim}, the kernels are built for dim={DIM}")
    res = _cu(binary_residuals, torch.uint8)
    if res.dim() != 2 or res.shape[1] != DIM * int(nbits) // 8:
        raise _lib.PlaidError(f"decompress_residuals: binary_residuals must be [n, {DIM * int(nbits) // 8}] for nbits={nbits}")
    codes = _cu(codes, torch.int32)
    n = res.shape[0]
    if codes.numel() != n:
        raise _lib.PlaidError("decompress_residuals: codes and binary_residuals disagree on the number of tokens")
    cent = _cu(centroids, torch.float16)
    W = build_weight_table(bucket_weights, reversed_bit_map, bucket_weight_combinations, nbits)
    out = torch.empty(n, DIM, device=res.device, dtype=torch.float16)
    if n:
        _lib.call("plaid_decompress_tokens_f16", _p(res), _p(codes), ctypes.c_int64(n), _p(W), _p(cent), cent.shape[0],
                  int(nbits), int(bool(normalize)), _p(out), _stream())
    return out


def packbits(bits):
    """Drop-in for ``ResidualCodec.packbits`` (residual.py:130,198; packbits.cu:10-57): flat u8 flags -> u8 [n/8],
    first flag in the most significant bit."""
    bits = _cu(bits, torch.uint8).reshape(-1)
    n = bits.numel()
    if n % 8:
        raise _lib.PlaidError(f"packbits: {n} flags is not a multiple of 8")
    out = torch.empty(max(n // 8, 1), device=bits.device, dtype=torch.uint8)
    if n:
        _lib.call("plaid_packbits", _p(bits), ctypes.c_int64(n), _p(out), _stream())
    return out[: n // 8]


def token_inv_norms(residuals, codes, weight_table, centroids_f16, nbits):
    """fp16 [n]: the scale factor 1 / max(||centroid + weights||, 1e-12) of every token, in the fused MaxSim kernel's
    own arithmetic (plaid_token_inv_norms); derived data of a loaded index."""
    res, codes = _cu(residuals, torch.uint8), _cu(codes, torch.int32)
    W, cent = _cu(weight_table, torch.float32), _cu(centroids_f16, torch.float16)
    n = codes.numel()
    out = torch.empty(max(n, 1), device=codes.device, dtype=torch.float16)
    if n:
        _lib.call("plaid_token_inv_norms", _p(res), _p(codes), ctypes.c_int64(n), _p(W), _p(cent), cent.shape[0], int(nbits),
                  _p(out), _stream())
    return out[:n]


def unpack_residual_codes(residuals, nbits, reversed_bit_map, lookup):
    """Integer parity tap: bucket index of every dimension, u8 [n, 128]."""
    residuals = _cu(residuals, torch.uint8)
    n = residuals.shape[0]
    out = torch.empty(max(n, 1), DIM, device=residuals.device, dtype=torch.uint8)
    rbm, lut = _cu(reversed_bit_map, torch.uint8), _cu(lookup, torch.uint8)   # named: must outlive the launch
    _lib.call("plaid_unpack_residual_codes", _p(residuals), n, int(nbits), _p(rbm), _p(lut), _p(out), _stream())
    return out[:n]


# --------------------------------------------------------------------------------------------- segmented ops
def segmented_maxsim(scores, lengths):
    """Drop-in for ``ColBERT.segmented_maxsim`` (colbert.py:311): scores f32 [T, nq], lengths i64
    [ndocs] -> f32 [ndocs] (zero-clamped per-token max, summed over query tokens)."""
    scores = _cu(scores, torch.float32)
    lengths = _cu(lengths, torch.int64)
    nd = lengths.numel()
    row_off = torch.zeros(nd + 1, device=scores.device, dtype=torch.int64)
    row_off[1:] = torch.cumsum(lengths, 0)
    out = torch.empty(max(nd, 1), device=scores.device, dtype=torch.float32)
    _lib.call("plaid_segmented_maxsim", _p(scores), scores.shape[1], _p(lengths), _p(row_off), nd, _p(out), _stream())
    return out[:nd]


def segmented_lookup(input, pids, lengths, offsets):
    """Drop-in for ``StridedTensor.segmented_lookup`` (strided_tensor.py:96): `lengths`/`offsets` are
    already indexed by pids; returns the packed rows."""
    del pids
    x = _cu(input)
    lengths = _cu(lengths, torch.int64)
    offsets = _cu(offsets, torch.int64)
    n = lengths.numel()
    out_off = torch.zeros(n + 1, device=x.device, dtype=torch.int64)
    out_off[1:] = torch.cumsum(lengths, 0)
    total = int(out_off[-1].item()) if n else 0
    row_shape = tuple(x.shape[1:])
    row_bytes = x.element_size()
    for s in row_shape:
        row_bytes *= s
    out = torch.empty((max(total, 1),) + row_shape, device=x.device, dtype=x.dtype)
    _lib.call("plaid_segmented_lookup", _p(x), row_bytes, _p(lengths), _p(offsets), _p(out_off), n, _p(out), _stream())
    return out[:total]


# --------------------------------------------------------------------------------------------- query prep
def prepare_queries(Q: torch.Tensor, remove_zero_rows: bool, Lq_pad: int | None = None, B_pad: int | None = None,
                    with_f16: bool = False):
    """Q f32 [B, Lq, 128] -> (Qb bf16 [B_pad, Lq_pad, 128], qlens i32 [B_pad]); with_f16 adds the fp16 twin
    (Qb, qlens, Qh)."""
    Q = _cu(Q, torch.float32)
    B, Lq, dim = Q.shape
    if dim != DIM:
        raise _lib.PlaidError(f"query dim {dim} != {DIM}")
    Lq_pad = Lq_pad or ((Lq + 31) // 32) * 32
    B_pad = B_pad or ((B + 3) // 4) * 4
    Qb = torch.empty(B_pad, Lq_pad, DIM, device=Q.device, dtype=torch.bfloat16)
    qlens = torch.empty(B_pad, device=Q.device, dtype=torch.int32)
    Qh = torch.empty(B_pad, Lq_pad, DIM, device=Q.device, dtype=torch.float16) if with_f16 else None
    _lib.call("plaid_prepare_queries", _p(Q), B, Lq, int(bool(remove_zero_rows)), B_pad, Lq_pad, _p(Qb),
              _p(Qh) if with_f16 else None, _p(qlens), _stream())
    return (Qb, qlens, Qh) if with_f16 else (Qb, qlens)


def to_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 (RNE) through the library (codebook / passage embeddings)."""
    x = _cu(x, torch.float32)
    out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    if x.numel():
        _lib.call("plaid_f32_to_bf16", _p(x), _p(out), x.numel(), _stream())
    return out
