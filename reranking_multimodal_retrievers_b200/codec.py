"""Index-build side of the residual codec on the GPU: the reference's
`ResidualCodec.compress_into_codes / compress / binarize` (CB/indexing/codecs/residual.py:169-222) and the IVF build
(CB/indexing/collection_indexer.py:393-431), so that an index can be created where it will be searched.

`compress_into_codes` is the centroid-scoring kernel run for its top-1 list only (no score table is written): 32
embeddings take the place of one query's tokens.  `compress` adds the fused residual / bucketize / bit-pack kernel.
Operands of the argmax are bf16 (tensor cores), so an embedding whose two best centroids are closer than bf16
resolution may land on the other one than the reference's fp32/fp16 matmul picks; everything after the code is
bit-exact given the code.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, ops
from .index import HostIndex, build_ivf
from .ops import NQ_MAX, _cu, _p, _stream

CELL_LISTS_PER_RANGE = 4      # PLAID_CELL_LISTS_PER_RANGE


def compress_into_codes(embs: torch.Tensor, centroids: torch.Tensor, batch_tokens: int = 1 << 18) -> torch.Tensor:
    """embs f32 [n, 128], centroids [C, 128] (any float dtype) -> i32 [n]: index of the centroid with the largest
    inner product (lowest id on ties, like torch's max)."""
    embs = _cu(embs, torch.float32)
    dev = embs.device
    C = centroids.shape[0]
    if C % 32:
        raise _lib.PlaidError(f"compress_into_codes: C={C} must be a multiple of 32")
    cent_bf16 = ops.to_bf16(_cu(centroids).to(dev).float())
    n = embs.shape[0]
    out = torch.empty(n, device=dev, dtype=torch.int32)
    wd = torch.zeros(1, device=dev, dtype=torch.int32)
    step = max(128, (batch_tokens // 128) * 128)
    for t0 in range(0, n, step):
        t1 = min(n, t0 + step)
        m = t1 - t0
        B = (m + NQ_MAX - 1) // NQ_MAX
        B_pad = ((B + 3) // 4) * 4
        Qb = torch.zeros(B_pad * NQ_MAX, 128, device=dev, dtype=torch.bfloat16)
        src = embs[t0:t1].contiguous()
        conv = ops.to_bf16(src)
        Qb[:m] = conv
        qlens = torch.zeros(B_pad, device=dev, dtype=torch.int32)
        qlens[:B] = NQ_MAX
        if m % NQ_MAX:
            qlens[B - 1] = m % NQ_MAX
        groups = B_pad // 4
        csplit = max(1, min((C + 255) // 256, 32, 148 // max(groups, 1)))
        nlists = CELL_LISTS_PER_RANGE * csplit
        cell_val = torch.empty(B_pad, NQ_MAX, nlists, 1, device=dev, dtype=torch.float32)
        cell_idx = torch.empty(B_pad, NQ_MAX, nlists, 1, device=dev, dtype=torch.int32)
        cells = torch.empty(B_pad, NQ_MAX, 1, device=dev, dtype=torch.int32)
        _lib.call("plaid_centroid_scores", _p(cent_bf16), C, _p(Qb), _p(qlens), B_pad, NQ_MAX, float("inf"), 1, csplit,
                  None, 0, None, _p(cell_val), _p(cell_idx), _p(wd), _stream())     # fp32 scores (no rounding before the argmax)
        _lib.call("plaid_merge_cells", _p(cell_val), _p(cell_idx), _p(qlens), B_pad, 1, nlists, _p(cells), _stream())
        out[t0:t1] = cells.reshape(-1)[:m]
    if int(wd.item()):
        raise _lib.PlaidError("compress_into_codes: a tcgen05 pipeline wait timed out (watchdog flag set)")
    return out


def compress_residuals(embs: torch.Tensor, codes: torch.Tensor, centroids_f16: torch.Tensor, bucket_cutoffs: torch.Tensor,
                       nbits: int) -> torch.Tensor:
    """u8 [n, 16*nbits]: residual.py:176-203 given the codes (bit-exact)."""
    embs = _cu(embs, torch.float32)
    dev = embs.device
    codes = _cu(codes, torch.int32)
    cent = _cu(centroids_f16.to(dev), torch.float16)
    cut = _cu(bucket_cutoffs.to(dev), torch.float32)
    if cut.numel() != (1 << nbits) - 1:
        raise _lib.PlaidError(f"compress_residuals: {cut.numel()} cutoffs for nbits={nbits}")
    n = embs.shape[0]
    out = torch.empty(n, 16 * nbits, device=dev, dtype=torch.uint8)
    bad = torch.zeros(1, device=dev, dtype=torch.int32)
    _lib.call("plaid_compress_residuals", _p(embs), _p(codes), _p(cent), _p(cut), ctypes.c_int64(n), cent.shape[0], int(nbits),
              _p(out), _p(bad), _stream())
    if int(bad.item()):
        raise _lib.PlaidError("compress_residuals: a code lies outside [0, C)")
    return out


def compress(embs: torch.Tensor, centroids_f16: torch.Tensor, bucket_cutoffs: torch.Tensor, nbits: int):
    """(codes i32 [n], residuals u8 [n, 16*nbits]) -- ResidualCodec.compress (residual.py:169-186)."""
    codes = compress_into_codes(embs, centroids_f16)
    return codes, compress_residuals(embs, codes, centroids_f16, bucket_cutoffs, nbits)


def build_index(embs: torch.Tensor, doclens: torch.Tensor, centroids_f16: torch.Tensor, bucket_cutoffs: torch.Tensor,
                bucket_weights: torch.Tensor, nbits: int) -> HostIndex:
    """Compress passage token embeddings [sum(doclens), 128] and build the inverted file on the device; the result
    feeds DeviceIndex / Searcher directly (tensors stay where `embs` lives)."""
    codes, residuals = compress(embs, centroids_f16, bucket_cutoffs, nbits)
    doclens = doclens.to(torch.int64)
    ivf, ivf_lengths = build_ivf(codes, doclens.to(codes.device), centroids_f16.shape[0])
    return HostIndex(centroids=centroids_f16.to(torch.float16), bucket_cutoffs=bucket_cutoffs.float(),
                     bucket_weights=bucket_weights.float(), codes=codes, residuals=residuals, doclens=doclens, ivf=ivf,
                     ivf_lengths=ivf_lengths, nbits=int(nbits))


class ResidualEmbeddings:
    """(codes i32 [n], residuals u8 [n, 16*nbits]) pair (CB/indexing/codecs/residual_embeddings.py:12-24)."""

    def __init__(self, codes, residuals):
        assert codes.size(0) == residuals.size(0), (codes.size(), residuals.size())
        assert codes.dim() == 1 and residuals.dim() == 2, (codes.size(), residuals.size())
        assert residuals.dtype == torch.uint8
        self.codes = codes.to(torch.int32)
        self.residuals = residuals


class ResidualCodec:
    """The reference's codec object (CB/indexing/codecs/residual.py:17-278), GPU branch only: same constructor,
    `compress_into_codes / lookup_centroids / compress / binarize / decompress`, and the two operators the reference
    binds as class attributes (residual.py:115,130)."""

    Embeddings = ResidualEmbeddings
    decompress_residuals = staticmethod(ops.codec_decompress_residuals)
    packbits = staticmethod(ops.packbits)

    def __init__(self, config, centroids, avg_residual=None, bucket_cutoffs=None, bucket_weights=None):
        if getattr(config, "total_visible_gpus", 1) == 0:
            raise RuntimeError("ResidualCodec: total_visible_gpus=0 selects the reference's CPU branch, which this "
                               "B200 implementation does not have")
        self.use_gpu = True
        self.dim, self.nbits = int(config.dim), int(config.nbits)
        self.centroids = _cu(centroids, torch.float16)
        self.avg_residual = avg_residual
        self.bucket_cutoffs = None if bucket_cutoffs is None else _cu(bucket_cutoffs, torch.float32)
        self.bucket_weights = None if bucket_weights is None else _cu(bucket_weights, torch.float16)
        from .index import codec_tables
        rbm, lut = codec_tables(self.nbits)
        self.reversed_bit_map, self.decompression_lookup_table = _cu(rbm), _cu(lut)
        self.arange_bits = torch.arange(0, self.nbits, device=self.centroids.device, dtype=torch.uint8)

    def compress_into_codes(self, embs, out_device=None):
        codes = compress_into_codes(embs.float(), self.centroids)
        return codes if out_device is None else codes.to(out_device)

    def lookup_centroids(self, codes, out_device=None):
        out = self.centroids[_cu(codes).long()]
        return out if out_device is None else out.to(out_device)

    def compress(self, embs):
        codes, residuals = compress(embs.float(), self.centroids, self.bucket_cutoffs, self.nbits)
        return ResidualEmbeddings(codes, residuals)

    def binarize(self, residuals):
        """residual.py:188-203 on the device: bucketize, nbits flags per value LSB first, packbits."""
        r = torch.bucketize(_cu(residuals).float(), self.bucket_cutoffs).to(torch.uint8)
        flags = (r.unsqueeze(-1) >> self.arange_bits) & 1
        packed = ResidualCodec.packbits(flags.contiguous().flatten())
        return packed.reshape(r.size(0), self.dim // 8 * self.nbits)

    def decompress(self, compressed_embs):
        """fp16 [n, dim], L2-normalised (residual.py:242-278, GPU branch) -- one launch, no 2^15-row batching."""
        return ResidualCodec.decompress_residuals(compressed_embs.residuals, self.bucket_weights, self.reversed_bit_map,
                                                  self.decompression_lookup_table, compressed_embs.codes,
                                                  self.centroids, self.dim, self.nbits, normalize=True)
