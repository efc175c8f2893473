"""The C restatement (oracle/plaid_oracle.c) against the reference's own compiled operators
(oracle/_ref, built from /root/reference by oracle/build_ref.py).  Skipped where oracle/_ref is absent."""
import os
import sys

import pytest
import torch

from plaid_test_helpers import ROOT, golden_oracle_index
from oracle import build_ref
from oracle import plaid_oracle as po

have_ref = all(os.path.exists(build_ref.ref_so_path(n)) for n in build_ref.SOURCES)
pytestmark = pytest.mark.skipif(not have_ref, reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def ref():
    return {n: build_ref.load(n) for n in build_ref.SOURCES}


def test_filter_pids_bit_exact(golden, ref):
    g = golden
    ix = golden_oracle_index(g)
    torch.set_num_threads(4)
    for b in (0, 1):
        S = torch.from_numpy(g[f"S_{b}"]).contiguous()
        idx = torch.from_numpy(g[f"idx_{b}"])
        cand = torch.from_numpy(g[f"cand_{b}"])
        ndocs = int(g["ndocs"])
        out = ref["filter_pids_cpp"].filter_pids_cpp(cand, S, ix.codes, ix.doclens, ix.offsets, idx, ndocs)
        mine = po.filter_pids(ix, cand, S, idx, ndocs)
        assert torch.equal(out, mine)


def test_decompress_bit_exact(golden, ref):
    g = golden
    ix = golden_oracle_index(g)
    pids = torch.from_numpy(g["stage2_0"])
    out = ref["decompress_residuals_cpp"].decompress_residuals_cpp(
        pids, ix.doclens, ix.offsets, ix.bucket_weights, ix.reversed_bit_map, ix.lookup, ix.residuals, ix.codes,
        ix.centroids, ix.dim, ix.nbits)
    assert torch.equal(out, po.decompress_residuals(ix, pids))


def test_segmented_maxsim(golden, ref):
    gen = torch.Generator().manual_seed(5)
    lengths = torch.tensor([3, 1, 17, 0, 9], dtype=torch.long)
    scores = torch.randn(int(lengths.sum()), 40, generator=gen)
    out = ref["segmented_maxsim_cpp"].segmented_maxsim_cpp(scores, lengths)
    mine = po.segmented_maxsim(scores, lengths)
    torch.testing.assert_close(out, mine, rtol=1e-6, atol=1e-6)
    assert mine[3] == 0  # empty passage: zero-initialised buffer


def test_segmented_lookup(golden, ref):
    g = golden
    ix = golden_oracle_index(g)
    pids = torch.tensor([5, 0, 17, 3], dtype=torch.long)
    lengths, offsets = ix.doclens[pids], ix.offsets[pids]
    out = ref["segmented_lookup_cpp"].segmented_lookup_cpp(ix.residuals, pids, lengths, offsets)
    total = int(lengths.sum())
    mine = torch.empty(total, ix.residuals.shape[1], dtype=torch.uint8)
    import ctypes
    po._lib().plaid_oracle_segmented_lookup(po._p(ix.residuals), ctypes.c_int64(ix.residuals.shape[1]), po._p(pids),
                                            ctypes.c_int(4), po._p(lengths.contiguous()), po._p(offsets.contiguous()),
                                            po._p(mine))
    assert torch.equal(out, mine)


def test_reference_cuda_operators_are_built_and_export_their_entry_points():
    """oracle/_ref also holds the reference's two CUDA codec operators (build_ref.py --gpu); they only RUN on the GPU box
    (tests/test_gpu_ops.py), here: the modules load and export the names residual.py:115,130 binds."""
    if not all(os.path.exists(build_ref.ref_so_path(n)) for n in build_ref.GPU_SOURCES):
        pytest.skip("reference CUDA operators not built (python oracle/build_ref.py --gpu)")
    assert callable(build_ref.load("decompress_residuals_gpu_cpp").decompress_residuals_cpp)
    assert callable(build_ref.load("packbits_gpu_cpp").packbits_cpp)


def test_reference_arm_runner_reproduces_the_unmodified_searcher(golden):
    """bench.py's `--impl reference` / cpu_baseline legs run oracle/ref_search.CpuSearcher: the reference's compiled
    operators under a restatement of IndexScorer.rank's Python glue (the reference's Python cannot travel to the GPU box).
    The golden vectors were recorded from the UNMODIFIED `Searcher._search_all_Q` (tests/golden/make_golden.py): the runner
    must return the same passages in the same order with the same scores -- what is timed as "the reference" computes
    what the reference computes."""
    import numpy as np
    from plaid_test_helpers import nonzero_rows
    from oracle.ref_search import CpuSearcher
    g = golden
    cs = CpuSearcher(golden_oracle_index(g), threads=4)
    assert cs.kind == "reference" and cs.cores == 4
    Q = torch.from_numpy(g["Q"])
    k = int(g["k"])
    toks = 0
    for b in range(Q.shape[0]):
        (pids, scores), t3 = cs.rank(nonzero_rows(Q[b]), int(g["ncells"]), float(g["threshold"]), int(g["ndocs"]))
        assert pids[:k] == g[f"rank_pids_{b}"].tolist()
        np.testing.assert_allclose(np.asarray(scores[:k], dtype=np.float32), g[f"rank_scores_{b}"], rtol=2e-6, atol=2e-5)
        assert sorted(pids) == sorted(g[f"stage2_{b}"].tolist())                 # the exact-scored set = stage-2 survivors
        toks += t3
    # T3 accounting of the arm: real tokens of the exact-scored passages (what bench.py divides by the time)
    doclens = torch.from_numpy(g["doclens"])
    assert toks == sum(int(doclens[torch.from_numpy(g[f"stage2_{b}"]).long()].sum()) for b in range(Q.shape[0]))


@pytest.mark.parametrize("nbits", [1, 8])
def test_restatement_equals_reference_operators_at_1_and_8_bits(ref, nbits):
    """The golden indexes are 2- and 4-bit; the reference also packs 1 and 8 bits per dimension (residual.py:47-89).
    On a synthetic index: the reference's compiled decompress_residuals_cpp == the C restatement bit for bit, and the whole
    CPU pipeline run through the reference's operators (the bench's reference arm) == the pure restatement."""
    from reranking_multimodal_retrievers_b200 import synthetic
    from oracle.ref_search import CpuSearcher
    sx = synthetic.make_synthetic_index(400, 8, 48, nbits, seed=20 + nbits, num_centroids=256, mode="codes")
    ix = po.OracleIndex(centroids=sx.centroids, bucket_weights=sx.bucket_weights, codes=sx.codes, residuals=sx.residuals,
                        doclens=sx.doclens, ivf=sx.ivf, ivf_lengths=sx.ivf_lengths, nbits=nbits)
    pids = torch.tensor([7, 0, 399, 123, 124], dtype=torch.int32)
    out = ref["decompress_residuals_cpp"].decompress_residuals_cpp(
        pids, ix.doclens, ix.offsets, ix.bucket_weights, ix.reversed_bit_map, ix.lookup, ix.residuals, ix.codes,
        ix.centroids, ix.dim, ix.nbits)
    assert torch.equal(out, po.decompress_residuals(ix, pids))
    Q = synthetic.make_queries(sx, 3, 40, seed=5)
    cs = CpuSearcher(ix, threads=2)
    assert cs.kind == "reference"
    for b in range(Q.shape[0]):
        (rp, rs), _ = cs.rank(Q[b], 2, 0.45, 128)
        r = po.rank(ix, Q[b], 2, 0.45, 128)
        assert rp == r["pids"].tolist()
        torch.testing.assert_close(torch.tensor(rs), r["scores"], rtol=2e-6, atol=2e-5)


def test_filter_pids_tie_order_matches_reference(golden, ref):
    """Stage scores tie massively in practice (every candidate without a surviving centroid scores the same): with a
    coarsely quantised table most candidates tie, and the restatement must keep the reference's (score, pid) order."""
    g = golden
    ix = golden_oracle_index(g)
    torch.set_num_threads(4)
    S = (torch.from_numpy(g["S_0"]) * 4).round() / 4                 # a handful of distinct values
    idx = S.max(-1).values >= 0.5
    cand = torch.arange(0, ix.doclens.numel(), 2, dtype=torch.int32)  # 500 candidates >= ndocs
    ndocs = int(g["ndocs"])
    assert cand.numel() >= ndocs
    out = ref["filter_pids_cpp"].filter_pids_cpp(cand, S.contiguous(), ix.codes, ix.doclens, ix.offsets, idx, ndocs)
    mine, (p1, s1, s2) = po.filter_pids(ix, cand, S.contiguous(), idx, ndocs, return_stages=True)
    assert torch.equal(out, mine)
    assert (s1[1:] == s1[:-1]).float().mean() > 0.3                  # the scenario really is tie-dominated
