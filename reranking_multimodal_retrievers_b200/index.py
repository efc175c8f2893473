"""PLAID index loading: the reference's on-disk format -> a GPU-resident, optionally pid-range-sharded
``DeviceIndex`` (reference loader CB/search/index_loader.py:13-85,
CB/indexing/codecs/residual_embeddings.py:27-52, CB/indexing/codecs/residual.py:134-150).

Layout in HBM (one shard):
  codes        i32 [NE]            nearest-centroid id per token, passage order
  residuals    u8  [NE + pad, 16*nbits]   packed bucket indices (16-byte aligned rows)
  offsets      i64 [N + 1]         exclusive prefix sum of doclens (strided_tensor_core.py:30-31)
  ivf_pids     i32 [sum lens]      per centroid, sorted unique LOCAL pids (CB/indexing/utils.py:8-53)
  ivf_offsets  i64 [C + 1]
  centroids_f16  f16 [C, 128]      exactly centroids.pt (gathered by the decompressor)
  centroids_bf16 bf16 [C, 128]     operand of the centroid-scoring contraction
  weight_table f32 [256, 8/nbits]  bucket_weights o lookup o reversed_bit_map (residual.py:54-89)
  inv_norms    f16 [NE]            1 / ||centroid[code] + residual weights|| per token (derived at load time)
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass

import torch

from . import ops


def codec_tables(nbits: int):
    """reversed_bit_map u8[256] and decompression_lookup_table u8[256, 8/nbits], restated from
    CB/indexing/codecs/residual.py:54-89 (bit-reverse every nbits-wide field; row r of the lookup
    table = the base-2^nbits digits of r, most significant first)."""
    mask, keys = (1 << nbits) - 1, 8 // nbits
    rbm = []
    for i in range(256):
        z = 0
        for f in range(keys):
            x = (i >> (8 - nbits * (f + 1))) & mask
            y = 0
            for bit in range(nbits):
                y |= ((x >> bit) & 1) << (nbits - 1 - bit)
            z = (z << nbits) | y
        rbm.append(z)
    lut = [[(r >> (nbits * (keys - 1 - l))) & mask for l in range(keys)] for r in range(256)]
    return torch.tensor(rbm, dtype=torch.uint8), torch.tensor(lut, dtype=torch.uint8)


@dataclass
class HostIndex:
    """CPU tensors of (a pid range of) an index, as the reference's IndexLoader holds them."""
    centroids: torch.Tensor        # f16 [C, 128]
    bucket_cutoffs: torch.Tensor
    bucket_weights: torch.Tensor   # f32 [2^nbits]
    codes: torch.Tensor            # i32 [NE]
    residuals: torch.Tensor        # u8 [NE, 16*nbits]
    doclens: torch.Tensor          # i64 [N]
    ivf: torch.Tensor | None       # i32 (LOCAL pids) or None -> rebuild from codes
    ivf_lengths: torch.Tensor | None
    nbits: int
    dim: int = 128
    pid_base: int = 0              # global pid of local passage 0
    num_passages_total: int | None = None
    config: dict | None = None
    residual_storage: torch.Tensor | None = None   # resident flat buffer behind `residuals` with 512 B of slack (device loads)


def build_ivf(codes: torch.Tensor, doclens: torch.Tensor, num_centroids: int, block_tokens: int = 1 << 27):
    """Per centroid the sorted unique pids owning a token with that code
    (CB/indexing/collection_indexer.py:393-431 + CB/indexing/utils.py:8-53).  Works in passage blocks of at most
    `block_tokens` tokens -- (code, pid) keys are sorted per block and the blocks' runs are scattered into place by
    a counting pass -- so that a 1.8 G-token shard needs a few GB of scratch, not a 64-bit sort of all its tokens."""
    dev = codes.device
    n = doclens.numel()
    doclens = doclens.to(dev, torch.int64)
    offsets = torch.zeros(n + 1, device=dev, dtype=torch.int64)
    offsets[1:] = torch.cumsum(doclens, 0)
    lengths = torch.zeros(num_centroids, device=dev, dtype=torch.int64)
    parts = []
    p0 = 0
    while p0 < n:
        target = offsets[p0] + block_tokens
        p1 = int(torch.searchsorted(offsets, target, right=True).item()) - 1
        p1 = min(n, max(p1, p0 + 1))
        e0, e1 = int(offsets[p0]), int(offsets[p1])
        tok2pid = torch.repeat_interleave(torch.arange(p0, p1, device=dev, dtype=torch.int64), doclens[p0:p1])
        key = torch.unique(codes[e0:e1].to(torch.int64) * n + tok2pid)       # sorted by (code, pid)
        del tok2pid
        c = torch.div(key, n, rounding_mode="floor")
        pid = (key - c * n).to(torch.int32)
        del key
        cnt = torch.bincount(c, minlength=num_centroids)
        lengths += cnt
        parts.append((c.to(torch.int32), pid, cnt))
        p0 = p1
    if len(parts) == 1:
        return parts[0][1].contiguous(), lengths.contiguous()
    total = int(lengths.sum().item())
    ivf = torch.empty(total, device=dev, dtype=torch.int32)
    base = torch.cumsum(lengths, 0) - lengths                                 # start of every centroid's list
    for c, pid, cnt in parts:                                                 # blocks are in pid order: lists stay sorted
        cl = c.long()
        start = torch.cumsum(cnt, 0) - cnt
        pos = base[cl] + (torch.arange(cl.numel(), device=dev, dtype=torch.int64) - start[cl])
        ivf[pos] = pid
        base += cnt
    return ivf.contiguous(), lengths.contiguous()


def _index_header(index_path: str, pid_range):
    with open(os.path.join(index_path, "metadata.json")) as f:
        meta = json.load(f)
    cfg = meta.get("config", {})
    nbits, dim = int(cfg["nbits"]), int(cfg.get("dim", 128))
    num_chunks = int(meta["num_chunks"])
    centroids = torch.load(os.path.join(index_path, "centroids.pt"), map_location="cpu")
    cutoffs, weights = torch.load(os.path.join(index_path, "buckets.pt"), map_location="cpu")
    chunk_doclens = []
    for i in range(num_chunks):
        with open(os.path.join(index_path, f"doclens.{i}.json")) as f:
            chunk_doclens.append(torch.tensor(json.load(f), dtype=torch.int64))
    n_total = sum(int(d.numel()) for d in chunk_doclens)
    p0, p1 = (0, n_total) if pid_range is None else (max(0, pid_range[0]), min(n_total, pid_range[1]))
    # the chunk files that overlap [p0, p1): (chunk index, first / last token of the slice inside the chunk, its doclens)
    todo, base = [], 0
    for i, dl in enumerate(chunk_doclens):
        c0, c1 = base, base + dl.numel()
        base = c1
        lo, hi = max(p0, c0), min(p1, c1)
        if lo >= hi:
            continue
        off = torch.cat((torch.zeros(1, dtype=torch.int64), torch.cumsum(dl, 0)))
        todo.append((i, int(off[lo - c0]), int(off[hi - c0]), dl[lo - c0:hi - c0]))
    return cfg, nbits, dim, centroids, cutoffs, weights, n_total, p0, p1, todo


def _load_chunk(index_path: str, i: int, e0: int, e1: int):
    codes = torch.load(os.path.join(index_path, f"{i}.codes.pt"), map_location="cpu")
    res = torch.load(os.path.join(index_path, f"{i}.residuals.pt"), map_location="cpu")
    return codes[e0:e1].to(torch.int32).contiguous(), res[e0:e1].contiguous()


def load_reference_index(index_path: str, pid_range: tuple[int, int] | None = None, device=None, workers: int = 4) -> HostIndex:
    """Read a reference-format index directory; with pid_range=(p0, p1) only that passage slice
    (only the chunk files that overlap it are opened).  The chunk files are read by `workers` threads.
    device=None: CPU tensors, as the reference's IndexLoader holds them (CB/search/index_loader.py:24-61).
    device=cuda: the codes / residuals go straight into their final device buffers -- every chunk is copied from a
    pinned staging buffer on a copy stream while the next chunk files are still being read -- and the result feeds
    DeviceIndex without another copy (the reference concatenates everything on the CPU first,
    CB/indexing/codecs/residual_embeddings.py:27-52)."""
    from concurrent.futures import ThreadPoolExecutor
    cfg, nbits, dim, centroids, cutoffs, weights, n_total, p0, p1, todo = _index_header(index_path, pid_range)
    pd = dim * nbits // 8
    doclens = torch.cat([t[3] for t in todo]) if todo else torch.zeros(0, dtype=torch.int64)
    total = sum(t[2] - t[1] for t in todo)
    storage = None
    with ThreadPoolExecutor(max_workers=max(1, workers)) as ex:
        futures = [ex.submit(_load_chunk, index_path, i, e0, e1) for i, e0, e1, _ in todo]
        if device is None:
            parts = [f.result() for f in futures]
            codes = torch.cat([c for c, _ in parts]) if parts else torch.zeros(0, dtype=torch.int32)
            residuals = torch.cat([r for _, r in parts]) if parts else torch.zeros(0, pd, dtype=torch.uint8)
        else:
            dev = torch.device(device)
            codes = torch.empty(total, device=dev, dtype=torch.int32)
            storage = torch.zeros(total * pd + 512, device=dev, dtype=torch.uint8)     # 512 B of slack, as DeviceIndex wants
            residuals = storage[: total * pd].view(total, pd)
            copy_stream = torch.cuda.Stream(device=dev)
            copy_stream.wait_stream(torch.cuda.current_stream(dev))
            biggest = max((e1 - e0 for _, e0, e1, _ in todo), default=0)
            stage = [(torch.empty(biggest, dtype=torch.int32).pin_memory(), torch.empty(biggest, pd, dtype=torch.uint8).pin_memory(),
                      torch.cuda.Event()) for _ in range(2 if todo else 0)]
            at = 0
            for j, f in enumerate(futures):
                c, r = f.result()
                n = c.numel()
                hc, hr, done = stage[j % 2]
                if j >= 2:
                    done.synchronize()                       # the copy that last used this staging pair has finished
                hc[:n].copy_(c)
                hr[:n].copy_(r)
                with torch.cuda.stream(copy_stream):
                    codes[at:at + n].copy_(hc[:n], non_blocking=True)
                    residuals[at:at + n].copy_(hr[:n], non_blocking=True)
                    done.record(copy_stream)
                at += n
            torch.cuda.current_stream(dev).wait_stream(copy_stream)
    ivf = ivf_lengths = None
    if pid_range is None or (p0 == 0 and p1 == n_total):
        ivf_path = os.path.join(index_path, "ivf.pid.pt")
        if os.path.exists(ivf_path):
            ivf, ivf_lengths = torch.load(ivf_path, map_location="cpu")
            ivf, ivf_lengths = ivf.to(torch.int32), ivf_lengths.to(torch.int64)
    host = HostIndex(centroids=centroids.half(), bucket_cutoffs=cutoffs.float(), bucket_weights=weights.float(),
                     codes=codes, residuals=residuals.contiguous(), doclens=doclens, ivf=ivf, ivf_lengths=ivf_lengths,
                     nbits=nbits, dim=dim, pid_base=p0, num_passages_total=n_total, config=cfg)
    host.residual_storage = storage
    return host


def shard_bounds(num_passages: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous pid range of `rank`: [rank*N/G, (rank+1)*N/G)."""
    return (num_passages * rank) // world_size, (num_passages * (rank + 1)) // world_size


def slice_host_index(ix, p0: int, p1: int) -> HostIndex:
    """pid-range slice of an in-memory index (HostIndex or synthetic.SyntheticIndex)."""
    doclens = ix.doclens.cpu() if ix.doclens.is_cuda else ix.doclens
    off = torch.cat((torch.zeros(1, dtype=torch.int64), torch.cumsum(doclens, 0)))
    e0, e1 = int(off[p0]), int(off[p1])
    return HostIndex(centroids=ix.centroids, bucket_cutoffs=ix.bucket_cutoffs, bucket_weights=ix.bucket_weights,
                     codes=ix.codes[e0:e1], residuals=ix.residuals[e0:e1], doclens=ix.doclens[p0:p1], ivf=None,
                     ivf_lengths=None, nbits=ix.nbits, dim=ix.dim, pid_base=p0 + getattr(ix, "pid_base", 0),
                     num_passages_total=int(doclens.numel()), config=getattr(ix, "config", None))


class DeviceIndex:
    """One shard of a PLAID index resident in HBM, in the layout the kernels read."""

    def __init__(self, host, device=None):
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.nbits, self.dim = int(host.nbits), int(host.dim)
        if self.dim != ops.DIM:
            raise ValueError(f"index dim {self.dim} != {ops.DIM}")
        self.pid_base = int(getattr(host, "pid_base", 0))
        self.config = getattr(host, "config", None)
        with torch.cuda.device(dev):
            self.doclens = host.doclens.to(dev, torch.int64).contiguous()
            self.num_passages = int(self.doclens.numel())
            self.offsets = torch.zeros(self.num_passages + 1, device=dev, dtype=torch.int64)
            self.offsets[1:] = torch.cumsum(self.doclens, 0)
            self.codes = host.codes.to(dev, torch.int32).contiguous()
            self.num_embeddings = int(self.codes.numel())
            pd = self.dim * self.nbits // 8
            # rows are read with 16-byte loads; keep 512 bytes of slack behind the last row
            res = getattr(host, "residual_storage", None)        # a resident flat buffer that already has the slack
            if not (res is not None and res.device == dev and res.dtype == torch.uint8 and res.dim() == 1
                    and res.numel() >= self.num_embeddings * pd + 512 and res.data_ptr() % 16 == 0):
                res = torch.zeros(self.num_embeddings * pd + 512, device=dev, dtype=torch.uint8)
                res[: self.num_embeddings * pd] = host.residuals.to(dev).reshape(-1)
            self._res_storage = res
            self.residuals = res[: self.num_embeddings * pd].view(self.num_embeddings, pd)
            self.centroids_f16 = host.centroids.to(dev, torch.float16).contiguous()
            self.num_centroids = int(self.centroids_f16.shape[0])
            if self.num_centroids % 128:
                raise ValueError("number of centroids must be a multiple of 128")
            ops.check_codes(self.codes, self.num_centroids)   # the kernels trust codes < C from here on
            self.centroids_f32 = self.centroids_f16.float()
            self.centroids_bf16 = ops.to_bf16(self.centroids_f32)
            self.bucket_weights = host.bucket_weights.to(dev, torch.float32).contiguous()
            rbm, lut = codec_tables(self.nbits)
            self.reversed_bit_map, self.lookup_table = rbm.to(dev), lut.to(dev)
            self.weight_table = ops.build_weight_table(self.bucket_weights, self.reversed_bit_map, self.lookup_table,
                                                       self.nbits)
            # per-token scale factors 1/||centroid + residual weights|| (fp16, 2 B/token; derived at load): the fused
            # decompress + MaxSim kernel multiplies by them instead of normalising every token it decodes
            self.inv_norms = ops.token_inv_norms(self.residuals, self.codes, self.weight_table, self.centroids_f16, self.nbits)
            if getattr(host, "ivf", None) is not None:
                ivf, ivf_lengths = host.ivf.to(dev, torch.int32), host.ivf_lengths.to(dev, torch.int64)
            else:
                ivf, ivf_lengths = build_ivf(self.codes, self.doclens, self.num_centroids)
            self.ivf_pids = ivf.contiguous()
            self.ivf_lengths = ivf_lengths.contiguous()
            self.ivf_offsets = torch.zeros(self.num_centroids + 1, device=dev, dtype=torch.int64)
            self.ivf_offsets[1:] = torch.cumsum(self.ivf_lengths, 0)
            self.max_doclen = int(self.doclens.max().item()) if self.num_passages else 0
            self.max_ivf_len = int(self.ivf_lengths.max().item()) if self.num_centroids else 0

    def bytes(self) -> int:
        ts = [self.doclens, self.offsets, self.codes, self._res_storage, self.centroids_f16, self.centroids_f32,
              self.centroids_bf16, self.ivf_pids, self.ivf_offsets]
        return sum(t.numel() * t.element_size() for t in ts)


def default_num_centroids(num_embeddings: int) -> int:
    """2^floor(log2(16*sqrt(NE)))  (CB/indexing/collection_indexer.py:98)."""
    return int(2 ** math.floor(math.log2(16 * math.sqrt(num_embeddings))))
