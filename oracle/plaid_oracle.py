"""oracle/plaid_oracle.py -- CPU restatement of the reference's PLAID search path.

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs; never from the product package.

Parity status: PINNED against (a) the reference's compiled operators in oracle/_ref and
(b) golden vectors recorded from the unmodified reference ``Searcher`` (tests/golden/).

The per-token loops live in plaid_oracle.c (plain C, exact float order); the glue around
them uses the same torch CPU ops the reference calls (matmul, topk, unique, sort,
F.normalize).  Paths cited are relative to /root/reference/third_party/ColBERT/colbert/.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c_oracle(force: bool = False) -> str:
    """gcc the C restatement into oracle/libplaid_oracle.so."""
    so = os.path.join(_HERE, "libplaid_oracle.so")
    src = os.path.join(_HERE, "plaid_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libplaid_oracle.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c_oracle())
        _LIB.plaid_oracle_decompress.restype = ctypes.c_int64
        _LIB.plaid_oracle_segmented_lookup.restype = ctypes.c_int64
        _LIB.plaid_oracle_select_top.restype = ctypes.c_int
    return _LIB


def _p(t: torch.Tensor):
    assert t.is_contiguous() and t.device.type == "cpu"
    return ctypes.c_void_p(t.data_ptr())


# --------------------------------------------------------------------------- codec tables
def codec_tables(nbits: int):
    """reversed_bit_map u8[256] and decompression_lookup_table u8[256, 8/nbits]
    (indexing/codecs/residual.py:54-89)."""
    rbm = torch.empty(256, dtype=torch.uint8)
    lut = torch.empty(256, 8 // nbits, dtype=torch.uint8)
    _lib().plaid_oracle_codec_tables(ctypes.c_int(nbits), _p(rbm), _p(lut))
    return rbm, lut


# --------------------------------------------------------------------------- index container
@dataclass
class OracleIndex:
    """The tensors IndexLoader (search/index_loader.py:13-61) holds, on CPU."""
    centroids: torch.Tensor        # f32 [C, dim]  (f16 on disk, .float() on CPU: residual.py:29)
    bucket_weights: torch.Tensor   # f32 [2^nbits]
    codes: torch.Tensor            # i32 [NE]
    residuals: torch.Tensor        # u8  [NE, dim*nbits/8]
    doclens: torch.Tensor          # i64 [N]
    ivf: torch.Tensor              # i32 [sum ivf_lengths]
    ivf_lengths: torch.Tensor      # i64 [C]
    nbits: int
    dim: int = 128
    bucket_cutoffs: torch.Tensor | None = None
    offsets: torch.Tensor = field(init=False)        # i64 [N+1] (strided_tensor_core.py:30-31)
    ivf_offsets: torch.Tensor = field(init=False)    # i64 [C+1]

    def __post_init__(self):
        self.centroids = self.centroids.float().contiguous()
        self.bucket_weights = self.bucket_weights.float().contiguous()
        self.codes = self.codes.to(torch.int32).contiguous()
        self.residuals = self.residuals.contiguous()
        self.doclens = self.doclens.long().contiguous()
        self.ivf = self.ivf.to(torch.int32).contiguous()
        self.ivf_lengths = self.ivf_lengths.long().contiguous()
        z = torch.zeros(1, dtype=torch.long)
        self.offsets = torch.cat((z, torch.cumsum(self.doclens, 0))).contiguous()
        self.ivf_offsets = torch.cat((z, torch.cumsum(self.ivf_lengths, 0))).contiguous()
        self.reversed_bit_map, self.lookup = codec_tables(self.nbits)


# --------------------------------------------------------------------------- candidate generation
def centroid_scores(centroids: torch.Tensor, Q: torch.Tensor) -> torch.Tensor:
    """S = centroids @ Q.T  -> [C, nq] fp32 (search/candidate_generation.py:13)."""
    return (centroids.float() @ Q.float().T).contiguous()


def get_cells(S: torch.Tensor, ncells: int) -> torch.Tensor:
    """Per query token the `ncells` best centroids, flattened + unique
    (candidate_generation.py:14-19).  torch.topk leaves exact ties unspecified; the
    restatement (and the CUDA path) break them by lowest centroid id."""
    C, nq = S.shape
    order = torch.argsort(S, dim=0, descending=True, stable=True)[:ncells]  # [ncells, nq]
    return torch.unique(order.permute(1, 0).flatten())


def cells_per_token(S: torch.Tensor, ncells: int) -> torch.Tensor:
    """[nq, ncells] centroid ids, best first, ties -> lowest id (layout of the CUDA output)."""
    return torch.argsort(S, dim=0, descending=True, stable=True)[:ncells].permute(1, 0).contiguous()


def candidate_pids(index: OracleIndex, cells: torch.Tensor) -> torch.Tensor:
    """ivf.lookup(cells) -> sort -> unique_consecutive (candidate_generation.py:31-37,57-60)."""
    parts = [index.ivf[index.ivf_offsets[c]:index.ivf_offsets[c + 1]] for c in cells.tolist()]
    if not parts:
        return torch.empty(0, dtype=torch.int32)
    pids = torch.cat(parts)
    return torch.unique_consecutive(pids.sort().values).to(torch.int32)


def centroid_mask(S: torch.Tensor, threshold: float) -> torch.Tensor:
    """idx = S.max(-1).values >= thr  (search/index_storage.py:115)."""
    return (S.max(-1).values >= threshold)


# --------------------------------------------------------------------------- filter_pids
def approx_scores(index: OracleIndex, pids, S, idx=None) -> torch.Tensor:
    """Per-document approximate score of filter_pids.cpp:27-72."""
    pids = pids.to(torch.int32).contiguous()
    S = S.float().contiguous()
    out = torch.empty(pids.numel(), dtype=torch.float32)
    idx8 = None if idx is None else idx.to(torch.uint8).contiguous()
    _lib().plaid_oracle_approx_scores(
        _p(pids), ctypes.c_int(pids.numel()), _p(S), ctypes.c_int(S.shape[1]), _p(index.codes),
        _p(index.doclens), _p(index.offsets), _p(idx8) if idx8 is not None else None, _p(out))
    return out


def select_top(pids, scores, keep):
    """(score, pid)-descending top-`keep` (filter_pids.cpp:108-123)."""
    pids = pids.to(torch.int32).contiguous()
    scores = scores.float().contiguous()
    op = torch.empty(min(keep, pids.numel()), dtype=torch.int32)
    os_ = torch.empty(min(keep, pids.numel()), dtype=torch.float32)
    n = _lib().plaid_oracle_select_top(_p(pids), _p(scores), ctypes.c_int(pids.numel()),
                                       ctypes.c_int(keep), _p(op), _p(os_))
    return op[:n], os_[:n]


def filter_pids(index: OracleIndex, pids, S, idx, ndocs, return_stages=False):
    """Two-stage centroid-only filter (filter_pids.cpp:126-164).  Returns the stage-2 pids
    (i32, (score,pid) descending); with return_stages also the stage-1 list and both score lists."""
    pids = pids.to(torch.int32).contiguous()
    S = S.float().contiguous()
    idx8 = idx.to(torch.uint8).contiguous()
    n = pids.numel()
    p1 = torch.empty(max(min(n, ndocs), 1), dtype=torch.int32)
    s1 = torch.empty_like(p1, dtype=torch.float32)
    p2 = torch.empty(max(min(n, ndocs // 4), 1), dtype=torch.int32)
    s2 = torch.empty_like(p2, dtype=torch.float32)
    n1, n2 = ctypes.c_int(0), ctypes.c_int(0)
    _lib().plaid_oracle_filter_pids(
        _p(pids), ctypes.c_int(n), _p(S), ctypes.c_int(S.shape[1]), _p(index.codes),
        _p(index.doclens), _p(index.offsets), _p(idx8), ctypes.c_int(ndocs),
        _p(p1), _p(s1), ctypes.byref(n1), _p(p2), _p(s2), ctypes.byref(n2))
    if return_stages:
        return p2[:n2.value], (p1[:n1.value], s1[:n1.value], s2[:n2.value])
    return p2[:n2.value]


# --------------------------------------------------------------------------- decompression
def unpack_residual_codes(index: OracleIndex, residuals: torch.Tensor) -> torch.Tensor:
    """Bucket index of every dimension: lut[rbm[byte]][l] (decompress_residuals.cpp:52-66)."""
    residuals = residuals.contiguous()
    out = torch.empty(residuals.shape[0], index.dim, dtype=torch.uint8)
    _lib().plaid_oracle_unpack_codes(_p(residuals), ctypes.c_int64(residuals.shape[0]),
                                     ctypes.c_int(index.dim), ctypes.c_int(index.nbits),
                                     _p(index.reversed_bit_map), _p(index.lookup), _p(out))
    return out


def decompress_residuals(index: OracleIndex, pids: torch.Tensor) -> torch.Tensor:
    """f32 [sum doclens[pids], dim]: bucket_weight + centroid (decompress_residuals.cpp:27-155)."""
    pids = pids.to(torch.int32).contiguous()
    total = int(index.doclens[pids.long()].sum())
    out = torch.empty(total, index.dim, dtype=torch.float32)
    n = _lib().plaid_oracle_decompress(
        _p(pids), ctypes.c_int(pids.numel()), _p(index.doclens), _p(index.offsets),
        _p(index.bucket_weights), _p(index.reversed_bit_map), _p(index.lookup),
        _p(index.residuals), _p(index.codes), _p(index.centroids), ctypes.c_int(index.dim),
        ctypes.c_int(index.nbits), _p(out))
    assert n == total
    return out


def normalize(D: torch.Tensor) -> torch.Tensor:
    """F.normalize(D.float(), p=2, dim=-1)  (index_storage.py:175)."""
    return torch.nn.functional.normalize(D.to(torch.float32), p=2, dim=-1)


# --------------------------------------------------------------------------- MaxSim
def segmented_maxsim(scores: torch.Tensor, lengths: torch.Tensor, return_max=False):
    """Zero-clamped per-document MaxSim over a packed [T, nq] score matrix
    (modeling/segmented_maxsim.cpp:22-93)."""
    scores = scores.float().contiguous()
    lengths = lengths.long().contiguous()
    nd, nq = lengths.numel(), scores.shape[1]
    out = torch.empty(nd, dtype=torch.float32)
    om = torch.empty(nd, nq, dtype=torch.float32) if return_max else None
    _lib().plaid_oracle_segmented_maxsim(_p(scores), _p(lengths), ctypes.c_int(nd), ctypes.c_int(nq),
                                         _p(out), _p(om) if om is not None else None)
    return (out, om) if return_max else out


def colbert_score_packed(Q: torch.Tensor, D_packed: torch.Tensor, D_lengths: torch.Tensor):
    """CPU branch of colbert_score_packed (modeling/colbert.py:289-311)."""
    Q = Q.squeeze(0) if Q.dim() == 3 else Q
    scores = D_packed @ Q.to(dtype=D_packed.dtype).T
    return segmented_maxsim(scores, D_lengths)


def colbert_score_reduce(scores_padded: torch.Tensor, D_mask: torch.Tensor):
    """-9999 fill, max over doc tokens, sum over query tokens (colbert.py:237-263, 'colbert'
    interaction).  Returns (scores, scores_padded) like the FLMR copy (flmr_utils.py:22-30)."""
    sp = scores_padded.clone()
    pad = ~D_mask.view(sp.size(0), sp.size(1)).bool()
    sp[pad] = -9999
    return sp.max(1).values.sum(-1), sp


def colbert_score(Q: torch.Tensor, D_padded: torch.Tensor, D_mask: torch.Tensor):
    """Padded MaxSim (colbert.py:268-286)."""
    scores = D_padded @ Q.to(dtype=D_padded.dtype).permute(0, 2, 1)
    return colbert_score_reduce(scores, D_mask)[0]


# --------------------------------------------------------------------------- full search
def search_defaults(k: int):
    """(ncells, centroid_score_threshold, ndocs) chosen by Searcher.dense_search (searcher.py:96-122)."""
    if k <= 100:
        return 2, 0.45, 1024
    return 4, 0.4, max(k * 4, 4096)


def remove_zero_rows(Q: torch.Tensor) -> torch.Tensor:
    """searcher.py:124-130 for one query [Lq, dim]."""
    return Q[torch.abs(Q).sum(dim=-1) > 0]


def rank(index: OracleIndex, Q: torch.Tensor, ncells: int, threshold: float, ndocs: int,
         query_maxlen: int = 32, S_override: torch.Tensor | None = None, taps: bool = False):
    """IndexScorer.rank, CPU branch (search/index_storage.py:67-98,100-184) for one query
    Q [Lq, dim].  S_override injects a centroid-score table (parity protocol).
    Final order is (score desc, pid desc)."""
    Qc = Q[:query_maxlen]
    S = centroid_scores(index.centroids, Qc) if S_override is None else S_override
    cells = get_cells(S, ncells)
    cand = candidate_pids(index, cells)
    idx = centroid_mask(S, threshold)
    p2, (p1, s1, s2) = filter_pids(index, cand, S, idx, ndocs, return_stages=True)
    D = normalize(decompress_residuals(index, p2))
    lens = index.doclens[p2.long()]
    scores = colbert_score_packed(Q, D, lens)
    order = np.lexsort((-p2.numpy().astype(np.int64), -scores.numpy().astype(np.float64)))
    order = torch.from_numpy(order)
    out = {"pids": p2[order], "scores": scores[order]}
    if taps:
        out.update(S=S, cells=cells, candidates=cand, idx=idx, stage1_pids=p1, stage1_scores=s1,
                   stage2_pids=p2, stage2_scores=s2, D=D, doclens=lens, scores_unsorted=scores)
    return out


def search_all(index: OracleIndex, Q: torch.Tensor, k: int, remove_zero_tensors: bool = True):
    """Searcher._search_all_Q / dense_search (searcher.py:80-136) for Q [B, Lq, dim]."""
    ncells, thr, ndocs = search_defaults(k)
    res = []
    for b in range(Q.shape[0]):
        q = remove_zero_rows(Q[b]) if remove_zero_tensors else Q[b]
        r = rank(index, q, ncells, thr, ndocs)
        res.append((r["pids"][:k], r["scores"][:k]))
    return res


# ---------------------------------------------------------------------------------------------------------------
# Index-build codec: numpy/torch restatement of ResidualCodec.compress (CPU branch),
# CB/indexing/codecs/residual.py:169-222.  Pinned by tests/golden/codec_nbits{2,4}.npz (recorded from the reference).
def codec_compress_into_codes(centroids: torch.Tensor, embs: torch.Tensor) -> torch.Tensor:
    """`(centroids @ batch.T).max(dim=0).indices` in fp32 (residual.py:204-222): first maximum = lowest id."""
    return (centroids.float() @ embs.float().T).max(dim=0).indices.to(torch.int32)


def codec_binarize(residuals: torch.Tensor, bucket_cutoffs: torch.Tensor, nbits: int) -> torch.Tensor:
    """residual.py:188-203: bucketize, nbits bits per dimension LSB first, packed MSB first."""
    import numpy as np
    b = torch.bucketize(residuals.float(), bucket_cutoffs.float()).to(torch.uint8)           # [n, dim]
    bits = (b.unsqueeze(-1) >> torch.arange(nbits, dtype=torch.uint8)) & 1                    # [n, dim, nbits]
    packed = np.packbits(np.asarray(bits.contiguous().flatten()))
    return torch.as_tensor(packed, dtype=torch.uint8).reshape(residuals.size(0), residuals.size(1) // 8 * nbits)


def codec_compress(centroids_f16: torch.Tensor, bucket_cutoffs: torch.Tensor, nbits: int, embs: torch.Tensor,
                   codes: torch.Tensor = None):
    """(codes, residual bytes) of residual.py:169-186; `codes` may be injected."""
    cent = centroids_f16.float()
    if codes is None:
        codes = codec_compress_into_codes(cent, embs)
    res = embs.float() - cent[codes.long()]
    return codes, codec_binarize(res, bucket_cutoffs, nbits)


def codec_decompress_gpu_form(bucket_weights, reversed_bit_map, lookup, binary_residuals, codes, centroids_f16,
                              normalize: bool = False):
    """Restatement of the reference's GPU-branch operator ResidualCodec.decompress_residuals
    (CB/indexing/codecs/decompress_residuals.cu:8-75): per element `out = half(bucket_weight); out += half(centroid)`,
    i.e. ONE half add (computed here as the fp32 sum of the two halves rounded to half: exact for |x| < 2, which
    unit-norm centroids plus residual weights satisfy).  normalize=True adds ResidualCodec.decompress's
    `F.normalize(x, p=2, dim=-1).half()` (residual.py:272-273) in the arithmetic torch's CUDA kernels use for half.
    Pinned on the GPU box: the reference kernel itself (compiled from /root/reference into oracle/_ref by
    oracle/build_ref.py --gpu) returns the same bits for nbits 1/2/4/8
    (tests/test_gpu_ops.py::test_gpu_form_codec_operators_equal_the_reference_cuda_kernels); on CPU it is pinned
    indirectly: the fp32 CPU operator on the same bytes (golden D_0) differs by one half rounding."""
    w = bucket_weights.half()[lookup[reversed_bit_map[binary_residuals.long()].long()].long()]
    w = w.reshape(binary_residuals.shape[0], -1)
    out = (w.float() + centroids_f16.half()[codes.long()].float()).half()
    if normalize:
        # F.normalize on a CUDA half tensor: the norm is accumulated in fp32 and ROUNDED TO HALF, clamp_min(1e-12) is
        # a no-op in half (1e-12 underflows to 0), then a half / half division (fp32 quotient rounded to half)
        nrm = out.float().pow(2).sum(-1, keepdim=True).sqrt().half().float()
        out = (out.float() / nrm).half()
    return out


def codec_packbits(flags: torch.Tensor) -> torch.Tensor:
    """ResidualCodec.packbits: the CPU branch's np.packbits (residual.py:200), which packbits.cu:10-57 reproduces
    (ballot over 32 flags, bit-reversed, bytes emitted most significant first)."""
    return torch.as_tensor(np.packbits(np.asarray(flags.contiguous().flatten().ne(0).to(torch.uint8))), dtype=torch.uint8)


# ---------------------------------------------------------------------------------------------------------------
# Training-time in-batch-negative scoring: restatement of FLMRModelForRetrieval.compute_ib_loss_new
# (src/models/flmr/models/flmr/modeling_flmr.py:1089-1125) with flmr_utils.colbert_score_reduce (flmr_utils.py:22-30).
# Pinned by tests/golden/ib_loss.npz, recorded by executing those two reference functions (tests/golden/make_ib_golden.py).
def ib_scores(Q: torch.Tensor, D: torch.Tensor, D_mask: torch.Tensor) -> torch.Tensor:
    """[B, B*n_docs] padded MaxSim of every query against every passage of the batch, fp32 (modeling_flmr.py:1098-1105)."""
    scores = (D.float().unsqueeze(0) @ Q.float().permute(0, 2, 1).unsqueeze(1)).flatten(0, 1)      # query-major
    pad = ~D_mask.repeat(Q.size(0), 1, 1).view(scores.size(0), scores.size(1)).bool()
    scores = scores.masked_fill(pad.unsqueeze(-1), -9999.0)
    return scores.max(1).values.sum(-1).reshape(Q.size(0), -1)


def ib_loss(Q: torch.Tensor, D: torch.Tensor, D_mask: torch.Tensor):
    """(loss, in_batch_scores, labels): cross entropy with the positive of query i at column i * (n_docs) (:1107-1123)."""
    s = ib_scores(Q, D, D_mask)
    step = D.shape[0] // Q.shape[0]
    labels = torch.arange(Q.shape[0], device=s.device) * step
    return torch.nn.functional.cross_entropy(s, labels), s, labels
