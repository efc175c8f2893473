"""Synthetic PLAID indexes and gold-planted FLMR-style queries.

Two generators, both device-agnostic torch code (run on cuda for the large bench shapes):

* ``mode="embed"`` (parity path): token embeddings are drawn around random unit centroids
  and then compressed the way the reference does -- nearest centroid by inner product,
  residual bucketised at its quantiles, bits emitted LSB-first and packed MSB-first
  (reference: indexing/codecs/residual.py:169-222, collection_indexer.py:296-313).
* ``mode="codes"`` (perf path): centroid assignments and residual bytes are drawn directly
  in code space; quantile buckets are equiprobable, so this is distribution-faithful and
  avoids the NE x C assignment GEMM.

``write_reference_format`` stores an index in the reference's on-disk layout (SURVEY.md
appendix B) so the loader -- and, in the authoring container, the unmodified reference
Searcher -- can read it.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass

import torch


@dataclass
class SyntheticIndex:
    centroids: torch.Tensor       # f16 [C, dim]
    bucket_cutoffs: torch.Tensor  # f32 [2^nbits - 1]
    bucket_weights: torch.Tensor  # f32 [2^nbits]
    avg_residual: torch.Tensor    # f32 [1]
    codes: torch.Tensor           # i32 [NE]
    residuals: torch.Tensor       # u8  [NE, dim*nbits/8]
    doclens: torch.Tensor         # i64 [N]
    ivf: torch.Tensor             # i32 [sum ivf_lengths]
    ivf_lengths: torch.Tensor     # i64 [C]
    nbits: int
    dim: int = 128

    @property
    def num_passages(self):
        return int(self.doclens.numel())

    @property
    def num_embeddings(self):
        return int(self.codes.numel())

    @property
    def num_centroids(self):
        return int(self.centroids.shape[0])

    def cpu(self):
        return SyntheticIndex(**{k: (v.cpu() if torch.is_tensor(v) else v) for k, v in self.__dict__.items()})


def default_num_centroids(num_embeddings: int) -> int:
    """2^floor(log2(16*sqrt(NE)))  (reference: indexing/collection_indexer.py:98)."""
    return int(2 ** math.floor(math.log2(16 * math.sqrt(num_embeddings))))


def build_ivf(codes: torch.Tensor, doclens: torch.Tensor, num_centroids: int):
    """Per centroid, the sorted unique pids owning a token with that code
    (reference: collection_indexer.py:393-431 + indexing/utils.py:8-53)."""
    dev = codes.device
    n = doclens.numel()
    tok2pid = torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int64), doclens.to(dev))
    key = codes.to(torch.int64) * n + tok2pid
    key = torch.unique(key)  # sorted
    ivf_codes = torch.div(key, n, rounding_mode="floor")
    ivf = (key - ivf_codes * n).to(torch.int32)
    ivf_lengths = torch.bincount(ivf_codes, minlength=num_centroids).to(torch.int64)
    return ivf.contiguous(), ivf_lengths.contiguous()


@dataclass
class CollectionShard:
    """A pid range of a large code-space collection, generated block by block straight into resident buffers
    (`residual_storage` already carries the slack DeviceIndex wants, so nothing is copied); the IVF is left to
    DeviceIndex (rebuilt from the shard's codes, as for any pid-range shard)."""
    centroids: torch.Tensor
    bucket_cutoffs: torch.Tensor
    bucket_weights: torch.Tensor
    codes: torch.Tensor
    residuals: torch.Tensor
    residual_storage: torch.Tensor
    doclens: torch.Tensor
    nbits: int
    pid_base: int
    num_passages_total: int
    dim: int = 128
    ivf: torch.Tensor | None = None
    ivf_lengths: torch.Tensor | None = None
    config: dict | None = None


def make_collection_shard(blocks, passages_per_block: int, doclen_lo: int, doclen_hi: int, nbits: int, num_centroids: int,
                          total_blocks: int, seed: int = 4000, device: str = "cuda", noise: float = 0.05, dim: int = 128):
    """Blocks `blocks` (consecutive indices) of a collection of total_blocks x passages_per_block passages in code space
    (mode="codes" of make_synthetic_index).  Block b is a function of (seed, b) alone and the codebook of `seed` alone,
    so the collection is the same however many ranks share it."""
    dev = torch.device(device)
    blocks = list(blocks)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    centroids = torch.nn.functional.normalize(torch.randn(num_centroids, dim, generator=g, device=dev), dim=-1).half()
    nb = 2 ** nbits
    sigma = noise * (1.0 - 1.0 / dim) ** 0.5
    q = torch.arange(nb, device=dev, dtype=torch.float32) / nb
    cutoffs = (sigma * _normal_icdf(q[1:])).float().cpu()
    weights = (sigma * _normal_icdf(q + 0.5 / nb)).float().cpu()
    doclens = []
    for b in blocks:
        g.manual_seed(seed + 1 + b)
        doclens.append(torch.randint(doclen_lo, doclen_hi + 1, (passages_per_block,), generator=g, device=dev, dtype=torch.int64))
    ne = [int(d.sum()) for d in doclens]
    total = sum(ne)
    pd = dim * nbits // 8
    codes = torch.empty(total, device=dev, dtype=torch.int32)
    storage = torch.zeros(total * pd + 512, device=dev, dtype=torch.uint8)
    residuals = storage[: total * pd].view(total, pd)
    e0 = 0
    for b, n_b in zip(blocks, ne):
        g.manual_seed(seed + 1000 + b)
        step = 1 << 26                                   # bounded temporaries
        for s0 in range(0, n_b, step):
            m = min(step, n_b - s0)
            codes[e0 + s0: e0 + s0 + m] = torch.randint(0, num_centroids, (m,), generator=g, device=dev, dtype=torch.int32)
            residuals[e0 + s0: e0 + s0 + m] = torch.randint(0, 256, (m, pd), generator=g, device=dev, dtype=torch.uint8)
        e0 += n_b
    return CollectionShard(centroids=centroids, bucket_cutoffs=cutoffs, bucket_weights=weights, codes=codes, residuals=residuals,
                           residual_storage=storage, doclens=torch.cat(doclens), nbits=nbits,
                           pid_base=blocks[0] * passages_per_block, num_passages_total=total_blocks * passages_per_block, dim=dim)


def binarize(bucket_idx: torch.Tensor, nbits: int) -> torch.Tensor:
    """u8 bucket indices [n, dim] -> packed residual bytes [n, dim*nbits/8].
    Bits are emitted LSB-first per value and packed MSB-first, exactly as
    residual.py:183-204 (``>> arange_bits``, ``& 1``, ``np.packbits``)."""
    n, dim = bucket_idx.shape
    ar = torch.arange(nbits, device=bucket_idx.device, dtype=torch.uint8)
    bits = (bucket_idx.unsqueeze(-1) >> ar) & 1                      # [n, dim, nbits]
    bits = bits.reshape(n, dim * nbits // 8, 8).to(torch.int32)
    w = torch.tensor([128, 64, 32, 16, 8, 4, 2, 1], device=bucket_idx.device, dtype=torch.int32)
    return (bits * w).sum(-1).to(torch.uint8).contiguous()


def _normal_icdf(p: torch.Tensor) -> torch.Tensor:
    return math.sqrt(2.0) * torch.erfinv(2 * p - 1)


def make_synthetic_index(num_passages: int, doclen_lo: int = 120, doclen_hi: int = 239, nbits: int = 2,
                         seed: int = 1234, num_centroids: int | None = None, mode: str = "codes",
                         device: str = "cpu", noise: float = 0.05, dim: int = 128) -> SyntheticIndex:
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    doclens = torch.randint(doclen_lo, doclen_hi + 1, (num_passages,), generator=g, device=dev, dtype=torch.int64)
    ne = int(doclens.sum())
    C = num_centroids or default_num_centroids(ne)
    centroids = torch.nn.functional.normalize(
        torch.randn(C, dim, generator=g, device=dev, dtype=torch.float32), dim=-1).half()
    nb = 2 ** nbits
    assign = torch.randint(0, C, (ne,), generator=g, device=dev, dtype=torch.int64)

    if mode == "codes":
        sigma = noise * (1.0 - 1.0 / dim) ** 0.5  # std of one residual coordinate (approx.)
        q = torch.arange(nb, device=dev, dtype=torch.float32) / nb
        bucket_cutoffs = (sigma * _normal_icdf(q[1:])).float()
        bucket_weights = (sigma * _normal_icdf(q + 0.5 / nb)).float()
        codes = assign.to(torch.int32)
        residuals = torch.randint(0, 256, (ne, dim * nbits // 8), generator=g, device=dev, dtype=torch.uint8)
        avg_residual = torch.tensor([sigma * math.sqrt(2 / math.pi)], dtype=torch.float32)
    elif mode == "embed":
        cf = centroids.float()
        codes = torch.empty(ne, dtype=torch.int32, device=dev)
        bucket_idx = torch.empty(ne, dim, dtype=torch.uint8, device=dev)
        chunk = max(1, (1 << 24) // C)
        res_sample, embs_chunks, budget = [], [], 1 << 22
        for s in range(0, ne, chunk):
            a = assign[s:s + chunk]
            e = torch.nn.functional.normalize(
                cf[a] + noise * torch.randn(a.numel(), dim, generator=g, device=dev), dim=-1)
            c = (cf @ e.T).max(dim=0).indices          # residual.py:207-222 (compress_into_codes)
            codes[s:s + chunk] = c.to(torch.int32)
            embs_chunks.append((s, e, c))
            if budget > 0:
                r = (e - cf[c]).flatten()[:budget]
                budget -= r.numel()
                res_sample.append(r)
        sample = torch.cat(res_sample).float()
        q = torch.arange(nb, device=dev, dtype=torch.float32) / nb
        bucket_cutoffs = sample.quantile(q[1:]).float()           # collection_indexer.py:308-313
        bucket_weights = sample.quantile(q + 0.5 / nb).float()
        avg_residual = sample.abs().mean().reshape(1).cpu()
        for s, e, c in embs_chunks:
            r = e - cf[c]
            bucket_idx[s:s + e.shape[0]] = torch.bucketize(r.float(), bucket_cutoffs).to(torch.uint8)
        residuals = binarize(bucket_idx, nbits)
    else:
        raise ValueError(mode)

    ivf, ivf_lengths = build_ivf(codes, doclens, C)
    return SyntheticIndex(centroids=centroids, bucket_cutoffs=bucket_cutoffs.cpu(),
                          bucket_weights=bucket_weights.cpu(), avg_residual=avg_residual,
                          codes=codes.contiguous(), residuals=residuals.contiguous(), doclens=doclens,
                          ivf=ivf, ivf_lengths=ivf_lengths, nbits=nbits, dim=dim)


def decompress_tokens(index: SyntheticIndex, token_ids: torch.Tensor) -> torch.Tensor:
    """Data-generation helper (NOT the search path): reconstruct a handful of tokens so
    queries can be planted near real passages.  Field l of raw byte x holds bit-reversed
    bucket index (x >> (8 - nbits*(l+1))) & mask (SURVEY.md 8a 'Residual bit layout')."""
    nbits, dim = index.nbits, index.dim
    dev = index.codes.device
    keys = 8 // nbits
    res = index.residuals[token_ids].to(torch.int32)                       # [n, pd]
    shifts = torch.tensor([8 - nbits * (l + 1) for l in range(keys)], device=dev, dtype=torch.int32)
    f = (res.unsqueeze(-1) >> shifts) & ((1 << nbits) - 1)                 # [n, pd, keys]
    b = torch.zeros_like(f)
    for i in range(nbits):
        b |= ((f >> i) & 1) << (nbits - 1 - i)
    w = index.bucket_weights.to(dev)[b.reshape(token_ids.numel(), dim).long()]
    return w + index.centroids[index.codes[token_ids].long()].float()


def make_queries(index: SyntheticIndex, num_queries: int, query_len: int = 64, seed: int = 99,
                 noise: float = 0.08, zero_rows: int = 0, return_gold: bool = False):
    """Gold-planted queries: each query copies `query_len` random tokens of one passage
    (decompressed + normalised) and perturbs them.  `zero_rows` trailing
    rows are zeroed to mimic PreFLMR's masked instruction tokens (searcher.py:124-130)."""
    dev = index.codes.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n = index.num_passages
    gold = torch.randint(0, n, (num_queries,), generator=g, device=dev)
    offsets = torch.cat((torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(index.doclens.to(dev), 0)))
    u = torch.rand(num_queries, query_len, generator=g, device=dev)
    tok = offsets[gold].unsqueeze(1) + (u * index.doclens.to(dev)[gold].unsqueeze(1)).long()
    D = torch.nn.functional.normalize(decompress_tokens(index, tok.flatten()), dim=-1)
    Q = torch.nn.functional.normalize(
        D + noise * torch.randn(D.shape, generator=g, device=dev), dim=-1).reshape(num_queries, query_len, index.dim)
    if zero_rows:
        Q[:, query_len - zero_rows:] = 0
    Q = Q.float().contiguous()
    return (Q, gold) if return_gold else Q


def write_reference_format(index: SyntheticIndex, index_path: str, chunk_passages: int = 25000,
                           query_maxlen: int = 32, extra_config: dict | None = None):
    """Write `index` as a reference-format PLAID index directory."""
    os.makedirs(index_path, exist_ok=True)
    ix = index.cpu()
    torch.save(ix.centroids.half(), os.path.join(index_path, "centroids.pt"))
    torch.save((ix.bucket_cutoffs, ix.bucket_weights), os.path.join(index_path, "buckets.pt"))
    torch.save(ix.avg_residual, os.path.join(index_path, "avg_residual.pt"))
    n = ix.num_passages
    offsets = torch.cat((torch.zeros(1, dtype=torch.int64), torch.cumsum(ix.doclens, 0)))
    nchunks = 0
    for ci, p0 in enumerate(range(0, n, chunk_passages)):
        p1 = min(n, p0 + chunk_passages)
        e0, e1 = int(offsets[p0]), int(offsets[p1])
        torch.save(ix.codes[e0:e1].clone(), os.path.join(index_path, f"{ci}.codes.pt"))
        torch.save(ix.residuals[e0:e1].clone(), os.path.join(index_path, f"{ci}.residuals.pt"))
        with open(os.path.join(index_path, f"doclens.{ci}.json"), "w") as f:
            json.dump(ix.doclens[p0:p1].tolist(), f)
        with open(os.path.join(index_path, f"{ci}.metadata.json"), "w") as f:
            json.dump({"passage_offset": p0, "num_passages": p1 - p0,
                       "num_embeddings": e1 - e0, "embedding_offset": e0}, f)
        nchunks += 1
    torch.save((ix.ivf, ix.ivf_lengths), os.path.join(index_path, "ivf.pid.pt"))
    config = {"dim": ix.dim, "nbits": ix.nbits, "query_maxlen": query_maxlen, "doc_maxlen": 512,
              "checkpoint": "synthetic", "collection": ["p"] * 3, "interaction": "colbert",
              "index_name": os.path.basename(index_path), "kmeans_niters": 4}
    config.update(extra_config or {})
    meta = {"config": config, "num_chunks": nchunks, "num_partitions": ix.num_centroids,
            "num_embeddings": ix.num_embeddings, "avg_doclen": ix.num_embeddings / max(n, 1)}
    with open(os.path.join(index_path, "metadata.json"), "w") as f:
        json.dump(meta, f, indent=2)
    return index_path
