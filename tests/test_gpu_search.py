"""GPU parity of the tensor-core kernels and of the whole search pipeline.

Protocol: float stages are compared within stated tolerances; every integer stage
(pruning mask, cells, candidate pids, stage-1 / stage-2 pids, top-k pids) must be BIT-EXACT when the
oracle is fed the same centroid-score table our kernel produced (injected inputs), because a bf16
contraction cannot reproduce an fp32 table bit for bit.
"""
import numpy as np
import pytest
import torch

from plaid_test_helpers import golden_host_index, golden_oracle_index, nonzero_rows
from oracle import plaid_oracle as po

pytestmark = pytest.mark.gpu

S_ABS_TOL = 4e-3        # bf16-rounded operands vs the fp32 reference table (measured 9.9e-4)
SCORE_REL_TOL = 1e-3    # north_star: 1e-3 relative on fp32-accumulated MaxSim


@pytest.fixture(scope="module")
def pkg():
    import reranking_multimodal_retrievers_b200 as p
    assert torch.cuda.is_available()
    return p


def _engine(g, fused=True, s_dtype=torch.float16):
    from reranking_multimodal_retrievers_b200.engine import SearchEngine
    from reranking_multimodal_retrievers_b200.index import DeviceIndex
    return SearchEngine(DeviceIndex(golden_host_index(g)), fused=fused, s_dtype=s_dtype)


S_DTYPES = [torch.float32, torch.float16]


@pytest.mark.parametrize("s_dtype", S_DTYPES)
def test_centroid_scores_tcgen05(pkg, golden, s_dtype):
    from reranking_multimodal_retrievers_b200 import ops
    g = golden
    eng = _engine(g, s_dtype=s_dtype)
    ix = golden_oracle_index(g)
    Q = torch.from_numpy(g["Q"])
    B = Q.shape[0]
    ncells, thr = int(g["ncells"]), float(g["threshold"])
    eng.search_batch(Q, k=int(g["k"]), ncells=ncells, centroid_score_threshold=thr, ndocs=int(g["ndocs"]),
                     remove_zero_rows=True, keep_taps=True)
    eng.check_flags()
    t = eng.last_taps
    cent_b = ix.centroids.bfloat16().float()
    for b in range(B):
        q = nonzero_rows(Q[b])
        nq = min(32, q.shape[0])
        assert int(t.qlens[b]) == q.shape[0]
        assert t.S.dtype == s_dtype
        S = t.S[b, :, :nq].float().cpu()
        # (1) same bf16-rounded operands, fp32 accumulate: only the summation order differs
        #     (fp16 storage: the stored value is that sum rounded to nearest fp16, |S| < 1 -> ulp <= 2^-11)
        S_ref_b = cent_b @ q[:nq].bfloat16().float().T
        if s_dtype == torch.float32:
            torch.testing.assert_close(S, S_ref_b, rtol=0, atol=2e-6)
        else:
            assert (S - S_ref_b).abs().max() <= 2 ** -12 + 2e-6
            assert (S != S_ref_b.half().float()).float().mean() < 1e-3
        # (2) against the fp32 reference table
        if f"S_{b}" in g:
            assert np.abs(S.numpy() - g[f"S_{b}"]).max() <= S_ABS_TOL
        # padded token lanes are exact zeros
        assert torch.count_nonzero(t.S[b, :, nq:].float()) == 0
        # pruning mask and cells are functions of OUR table: bit-exact
        idx = ops.unpack_idx_bits(t.idx_bits[b], S.shape[0]).cpu()
        assert torch.equal(idx, po.centroid_mask(S, thr))
        cells = t.cells[b, :nq].cpu()
        assert torch.equal(cells.long(), po.cells_per_token(S, ncells))
        assert torch.all(t.cells[b, nq:] == -1)


@pytest.mark.parametrize("s_dtype", S_DTYPES)
def test_pipeline_stagewise_bit_exact_with_injected_table(pkg, golden, s_dtype):
    from reranking_multimodal_retrievers_b200 import ops
    g = golden
    eng = _engine(g, fused=False, s_dtype=s_dtype)   # the unfused kernel pair materialises D, which this test inspects
    ix = golden_oracle_index(g)
    Q = torch.from_numpy(g["Q"])
    B = Q.shape[0]
    ncells, thr, ndocs, k = int(g["ncells"]), float(g["threshold"]), int(g["ndocs"]), int(g["k"])
    pids, scores, counts = eng.search_batch(Q, k=k, ncells=ncells, centroid_score_threshold=thr, ndocs=ndocs,
                                            remove_zero_rows=True, keep_taps=True)
    eng.check_flags()
    t = eng.last_taps
    # the fused kernel (decompression feeding the tensor cores through shared memory) builds the very same
    # fp16 operand tiles, so its scores and ranking are bit-identical to the unfused pair's
    engf = _engine(g, fused=True, s_dtype=s_dtype)
    pf, sf, cf = engf.search_batch(Q, k=k, ncells=ncells, centroid_score_threshold=thr, ndocs=ndocs,
                                   remove_zero_rows=True, keep_taps=True)
    engf.check_flags()
    assert torch.equal(pf, pids) and torch.equal(sf, scores) and torch.equal(cf, counts)
    for b in range(B):
        n2 = int(t.stage2_counts[b])
        assert torch.equal(engf.last_taps.scores[b, :n2], t.scores[b, :n2])
    for b in range(B):
        q = nonzero_rows(Q[b])
        nq = min(32, q.shape[0])
        S = t.S[b, :, :nq].float().cpu().contiguous()
        r = po.rank(ix, q, ncells, thr, ndocs, S_override=S, taps=True)
        nc = int(t.cand_counts[b])
        assert torch.equal(t.cand_pids[b, :nc].cpu(), r["candidates"])                     # sorted unique pids
        n1, n2 = int(t.stage1_counts[b]), int(t.stage2_counts[b])
        assert torch.equal(t.stage1_pids[b, :n1].cpu(), r["stage1_pids"])
        assert torch.equal(t.stage1_scores[b, :n1].cpu(), r["stage1_scores"])
        assert torch.equal(t.stage2_pids[b, :n2].cpu(), r["stage2_pids"])
        assert torch.equal(t.stage2_scores[b, :n2].cpu(), r["stage2_scores"])
        # decompressed + normalised passages (fp16 arithmetic, as the reference's GPU branch): a few fp16 ulps
        # from the oracle's fp32 rows.
        # In D every passage starts on a 32-token boundary; the rows in between are zero.
        lens = ix.doclens[r["stage2_pids"].long()]
        to = t.tok_offsets[b, :n2 + 1].cpu().long()
        assert torch.equal(to[1:] - to[:-1], (lens + 31) // 32 * 32) and int(to[0]) == 0
        Dq = t.D[b * t.tok_stride: b * t.tok_stride + int(to[-1])].float().cpu()
        rows = torch.cat([torch.arange(int(o), int(o) + int(l)) for o, l in zip(to[:-1], lens)])
        pad = torch.ones(Dq.shape[0], dtype=torch.bool)
        pad[rows] = False
        assert torch.count_nonzero(Dq[pad]) == 0
        D = Dq[rows]
        assert (D - r["D"]).abs().max() <= 2 ** -10
        assert (D - r["D"]).abs().mean() <= 1e-4
        assert ((D.norm(dim=-1) - 1).abs() <= 1e-3).all()
        # exact MaxSim: (a) same fp16 operands -> only summation order differs; (b) fp32 oracle within 1e-3
        sc = t.scores[b, :n2].cpu()
        ref_same = po.colbert_score_packed(q.half().float(), D, lens)
        torch.testing.assert_close(sc, ref_same, rtol=2e-5, atol=2e-5)
        assert ((sc - r["scores_unsorted"]).abs() <= SCORE_REL_TOL * r["scores_unsorted"].abs() + 1e-6).all()
        # final order: (score desc, pid desc) of our own scores, exactly
        m = int(counts[b])
        assert m == min(k, n2)
        rp, rs = po.select_top(r["stage2_pids"], sc, k)
        assert torch.equal(pids[b, :m].cpu(), rp) and torch.equal(scores[b, :m].cpu(), rs)
        # against the reference's own ranking (recorded): the planted passage leads, scores agree
        assert int(pids[b, 0]) == int(g[f"rank_pids_{b}"][0])
        gs = torch.from_numpy(g[f"rank_scores_{b}"])
        assert ((scores[b, :m].cpu() - gs[:m]).abs() <= SCORE_REL_TOL * gs[:m].abs() + 1e-6).all()


@pytest.mark.parametrize("nbits,Lq,lo,hi,zero_rows", [(4, 320, 128, 512, 0), (2, 64, 1, 40, 0), (1, 96, 1, 70, 9),
                                                       (8, 33, 5, 300, 0), (2, 200, 100, 239, 60)])
def test_search_all_bit_widths_and_query_lengths(pkg, nbits, Lq, lo, hi, zero_rows):
    """Fused decompress+MaxSim over 1/2/4/8-bit residuals, 1-3 query m-tiles (PreFLMR's 320 tokens),
    passages of 1..512 tokens, masked (all-zero) query rows -- against the oracle with our table injected."""
    from reranking_multimodal_retrievers_b200.engine import SearchEngine
    from reranking_multimodal_retrievers_b200.index import DeviceIndex
    from reranking_multimodal_retrievers_b200 import synthetic
    sx = synthetic.make_synthetic_index(1500, lo, hi, nbits, seed=7 * nbits + Lq, num_centroids=512, mode="codes")
    Q = synthetic.make_queries(sx, 6, Lq, seed=3, zero_rows=zero_rows)
    ix = po.OracleIndex(centroids=sx.centroids, bucket_weights=sx.bucket_weights, codes=sx.codes, residuals=sx.residuals,
                        doclens=sx.doclens, ivf=sx.ivf, ivf_lengths=sx.ivf_lengths, nbits=nbits)
    k, ndocs, ncells, thr = 20, 256, 2, 0.45
    res = {}
    for fused in (True, False):
        eng = SearchEngine(DeviceIndex(sx), fused=fused)
        res[fused] = eng.search_batch(Q, k=k, ncells=ncells, centroid_score_threshold=thr, ndocs=ndocs,
                                      remove_zero_rows=True, keep_taps=True)
        eng.check_flags()
        t = eng.last_taps
    for a, c in zip(res[True], res[False]):
        assert torch.equal(a, c)
    pids, scores, counts = res[True]
    for b in range(Q.shape[0]):
        q = nonzero_rows(Q[b])
        nq = min(32, q.shape[0])
        S = t.S[b, :, :nq].float().cpu().contiguous()
        r = po.rank(ix, q, ncells, thr, ndocs, S_override=S, taps=True)
        n2 = int(t.stage2_counts[b])
        assert torch.equal(t.stage2_pids[b, :n2].cpu(), r["stage2_pids"])
        sc = t.scores[b, :n2].cpu()
        assert ((sc - r["scores_unsorted"]).abs() <= SCORE_REL_TOL * r["scores_unsorted"].abs() + 1e-5).all()
        m = int(counts[b])
        rp, rs = po.select_top(r["stage2_pids"], sc, k)
        assert torch.equal(pids[b, :m].cpu(), rp) and torch.equal(scores[b, :m].cpu(), rs)
        assert int(pids[b, 0]) == int(r["pids"][0])


@pytest.mark.parametrize("thr,cap_s,cap_p", [(0.45, 4096, 65536), (0.45, 4096, 64), (0.45, 8, 65536), (0.1, 4096, 65536),
                                             (0.25, 4096, 65536), (-1.0, 4096, 65536)])
def test_stage1_ivf_route_equals_token_scan(pkg, golden, thr, cap_s, cap_p):
    """Stage 1 through the inverted file vs the token scan: identical lists and scores, whichever route the device
    picks per query (sparse masks -> IVF pairs; dense masks / workspace overflow -> scan)."""
    from reranking_multimodal_retrievers_b200.engine import SearchEngine
    from reranking_multimodal_retrievers_b200.index import DeviceIndex
    g = golden
    ix = DeviceIndex(golden_host_index(g))
    Q = torch.from_numpy(g["Q"])
    kw = dict(k=int(g["k"]), ncells=int(g["ncells"]), centroid_score_threshold=thr, ndocs=int(g["ndocs"]),
              remove_zero_rows=True, keep_taps=True)
    a = SearchEngine(ix, ivf_stage1=True)
    a.cap_s, a.cap_p = cap_s, cap_p
    b = SearchEngine(ix, ivf_stage1=False)
    ra, rb = a.search_batch(Q, **kw), b.search_batch(Q, **kw)
    a.check_flags(); b.check_flags()
    ta, tb = a.last_taps, b.last_taps
    for x, y in zip(ra, rb):
        assert torch.equal(x, y)
    B = Q.shape[0]
    assert torch.equal(ta.stage1_counts[:B], tb.stage1_counts[:B])
    for q in range(B):
        n1 = int(ta.stage1_counts[q])
        assert torch.equal(ta.stage1_pids[q, :n1], tb.stage1_pids[q, :n1])
        assert torch.equal(ta.stage1_scores[q, :n1], tb.stage1_scores[q, :n1])
    used_scan = a._ws["ivf_meta"][:B, 2].cpu()
    if thr == 0.45 and cap_s == 4096 and cap_p == 65536:
        assert int(used_scan.sum()) == 0            # sparse masks: every query went through the IVF
    if cap_p == 64 or cap_s == 8:
        assert int(used_scan.sum()) == B            # workspace too small: every query fell back to the scan


@pytest.mark.parametrize("range_slots", [100, 1024])
def test_stage1_ivf_route_in_slot_ranges(pkg, golden, range_slots):
    """Candidate lists longer than the shared-memory bins of the inverted-file stage 1 (shards of millions of passages)
    are sorted and reduced in slot ranges; forced here on the small index through the test hook: same lists, same bits."""
    from reranking_multimodal_retrievers_b200 import _lib
    from reranking_multimodal_retrievers_b200.engine import SearchEngine
    from reranking_multimodal_retrievers_b200.index import DeviceIndex
    g = golden
    ix = DeviceIndex(golden_host_index(g))
    Q = torch.from_numpy(g["Q"])
    kw = dict(k=int(g["k"]), ncells=int(g["ncells"]), centroid_score_threshold=0.45, ndocs=int(g["ndocs"]),
              remove_zero_rows=True, keep_taps=True)
    a, b = SearchEngine(ix, ivf_stage1=True), SearchEngine(ix, ivf_stage1=False)
    prev = _lib.lib().plaid_set_ivf_range_slots(range_slots)
    try:
        ra = a.search_batch(Q, **kw)
        torch.cuda.synchronize()
    finally:
        _lib.lib().plaid_set_ivf_range_slots(prev)
    rb = b.search_batch(Q, **kw)
    a.check_flags(); b.check_flags()
    B = Q.shape[0]
    assert int(a._ws["cand_counts"][:B].max()) > 2 * range_slots or range_slots > 100   # several ranges per query
    assert int(a._ws["ivf_meta"][:B, 2].sum()) == 0                                      # nobody took the scan
    for x, y in zip(ra, rb):
        assert torch.equal(x, y)
    ta, tb = a.last_taps, b.last_taps
    assert torch.equal(ta.stage1_counts[:B], tb.stage1_counts[:B])
    for q in range(B):
        n1 = int(ta.stage1_counts[q])
        assert torch.equal(ta.stage1_pids[q, :n1], tb.stage1_pids[q, :n1])
        assert torch.equal(ta.stage1_scores[q, :n1], tb.stage1_scores[q, :n1])


def test_colbert_score_padded_vs_reference(pkg, golden):
    g = golden
    Q, D, mask = (torch.from_numpy(g[n]) for n in ("cs_Q", "cs_D", "cs_mask"))
    out = pkg.colbert_score(Q, D, mask)
    ref = torch.from_numpy(g["cs_scores"])
    assert out.dtype == torch.float32 and out.shape == ref.shape
    # 1e-3 relative to the magnitude of what is summed: sum_k |max_t <q_k, d_t>| (equal to |score| whenever
    # the per-token maxima share a sign, which the packed path's clamp at 0 guarantees; these random
    # unit vectors produce signed maxima that cancel)
    full = D @ Q.permute(0, 2, 1)
    mag = po.colbert_score_reduce(full, mask)[1].max(1).values.abs().sum(-1)
    assert ((out.cpu() - ref).abs() <= SCORE_REL_TOL * mag + 1e-5).all()
    # same bf16 operands: tight
    ref_b = po.colbert_score(Q.bfloat16().float(), D.bfloat16().float(), mask)
    torch.testing.assert_close(out.cpu(), ref_b, rtol=2e-5, atol=2e-4)
    # FLMR form: also returns the masked similarity matrix [n, Ld, Lq]
    s2, raw = pkg.flmr_colbert_score(Q, D, mask)
    assert torch.equal(s2, out)
    full = (D.bfloat16().float() @ Q.bfloat16().float().permute(0, 2, 1))
    _, raw_ref = po.colbert_score_reduce(full, mask)
    torch.testing.assert_close(raw.cpu(), raw_ref, rtol=0, atol=2e-5)
    # one query block per passage (Q.size(0) == n, colbert.py:276-281)
    Qn = Q.repeat(D.shape[0], 1, 1).contiguous()
    torch.testing.assert_close(pkg.colbert_score(Qn, D, mask), out, rtol=0, atol=0)


@pytest.mark.parametrize("Lq,Ld,n,dpq", [(32, 7, 5, 5), (64, 180, 300, 100), (160, 33, 40, 8), (320, 97, 50, 25),
                                         (500, 16, 9, 3)])
def test_colbert_score_padded_shapes(pkg, Lq, Ld, n, dpq):
    """Ragged masks, several m-tiles (Lq up to 512), passages spanning tile boundaries, short groups."""
    g = torch.Generator().manual_seed(Lq * 1000 + Ld)
    nQ = (n + dpq - 1) // dpq
    Q = torch.nn.functional.normalize(torch.randn(nQ, Lq, 128, generator=g), dim=-1)
    D = torch.nn.functional.normalize(torch.randn(n, Ld, 128, generator=g), dim=-1)
    lens = torch.randint(1, Ld + 1, (n,), generator=g)
    lens[0] = Ld
    mask = torch.arange(Ld).unsqueeze(0) < lens.unsqueeze(1)
    if n > 3:
        mask[3] = False                                   # fully padded passage
    out = pkg.colbert_score(Q, D, mask, docs_per_query=dpq).cpu()
    Qd = Q.bfloat16().float().repeat_interleave(dpq, dim=0)[:n]
    ref = po.colbert_score(Qd, D.bfloat16().float(), mask)
    torch.testing.assert_close(out, ref, rtol=2e-5, atol=2e-3)


@pytest.mark.parametrize("Lq,lens", [(64, [3, 1, 170, 0, 9, 64, 128, 129, 0]), (320, [40, 0, 0, 300, 17]),
                                     (32, [0]), (45, [513, 2])])
def test_colbert_score_packed_edge_cases(pkg, Lq, lens):
    """Empty passages, passages longer than a tile, passage ends on tile boundaries."""
    g = torch.Generator().manual_seed(Lq + len(lens))
    lengths = torch.tensor(lens, dtype=torch.long)
    T = int(lengths.sum())
    Q = torch.nn.functional.normalize(torch.randn(1, Lq, 128, generator=g), dim=-1)
    D = torch.nn.functional.normalize(torch.randn(max(T, 1), 128, generator=g), dim=-1)[:T]
    out = pkg.colbert_score_packed(Q, D, lengths).cpu()
    ref = po.colbert_score_packed(Q.bfloat16().float(), D.bfloat16().float(), lengths)
    torch.testing.assert_close(out, ref, rtol=2e-5, atol=2e-5)
    for i, l in enumerate(lens):
        if l == 0:
            assert float(out[i]) == 0.0                  # zero-initialised max buffer (segmented_maxsim.cpp:58-59)


def test_searcher_dropin_on_reference_format_index(pkg, golden, tmp_path):
    """create_searcher / search_custom_collection read a reference-format directory and return the
    same Ranking structure FLMR_base_executor consumes."""
    from reranking_multimodal_retrievers_b200.synthetic import SyntheticIndex, write_reference_format
    g = golden
    nbits = int(g["nbits"])
    sx = SyntheticIndex(
        centroids=torch.from_numpy(g["centroids"]), bucket_cutoffs=torch.from_numpy(g["bucket_cutoffs"]),
        bucket_weights=torch.from_numpy(g["bucket_weights"]), avg_residual=torch.zeros(1),
        codes=torch.from_numpy(g["codes"]), residuals=torch.from_numpy(g["residuals"]),
        doclens=torch.from_numpy(g["doclens"]), ivf=torch.from_numpy(g["ivf"]),
        ivf_lengths=torch.from_numpy(g["ivf_lengths"]), nbits=nbits)
    path = tmp_path / "exp" / "indexes" / f"golden.nbits={nbits}"
    write_reference_format(sx, str(path), chunk_passages=300)       # several chunk files
    searcher = pkg.create_searcher(str(tmp_path), "exp", "golden", use_gpu=True, nbits=nbits)
    searcher.configure(ndocs=int(g["ndocs"]))
    Q = torch.from_numpy(g["Q"])
    k = int(g["k"])
    ranking = pkg.search_custom_collection(searcher, {i: f"q{i}" for i in range(Q.shape[0])}, Q,
                                           num_document_to_retrieve=k, remove_zero_tensors=True)
    rk = ranking.todict()
    for b in range(Q.shape[0]):
        rows = rk[b]
        assert len(rows) == k and [r[1] for r in rows] == list(range(1, k + 1))
        assert rows[0][0] == int(g[f"rank_pids_{b}"][0])
        gs = g[f"rank_scores_{b}"]
        assert np.allclose([r[2] for r in rows], gs, rtol=SCORE_REL_TOL, atol=1e-6)
        # single-query API gives the same answer as the batch path
        p1, ranks, s1 = searcher.dense_search(Q[b:b + 1], k=k, remove_zero_tensors=True)
        assert p1 == [r[0] for r in rows] and np.allclose(s1, [r[2] for r in rows], rtol=0, atol=0)
    # IndexScorer.rank with a filter_fn (reference signature): drop the best passage, it must vanish
    best = rk[0][0][0]
    q0 = nonzero_rows(Q[0]).unsqueeze(0)
    p, s = searcher.ranker.rank(searcher.config, q0, filter_fn=lambda pids: pids[pids != best])
    assert best not in p and len(p) > 0 and s == sorted(s, reverse=True)


def test_sharded_search_equals_reference_per_shard_merge(pkg, golden):
    """Oracle (A) of SURVEY 8e: run each pid-range shard separately, merge by (score, pid).  All G
    shards are emulated on one GPU (one engine per shard), the merge goes through plaid_merge_topk."""
    from reranking_multimodal_retrievers_b200 import sharded
    from reranking_multimodal_retrievers_b200.engine import SearchEngine
    from reranking_multimodal_retrievers_b200.index import DeviceIndex, shard_bounds, slice_host_index
    g = golden
    host = golden_host_index(g)
    Q = torch.from_numpy(g["Q"])
    G, k = 4, int(g["k"])
    N = host.doclens.numel()
    kw = dict(k=k, ncells=int(g["ncells"]), centroid_score_threshold=float(g["threshold"]), ndocs=int(g["ndocs"]),
              remove_zero_rows=True)
    lists = []
    for r in range(G):
        p0, p1 = shard_bounds(N, G, r)
        eng = SearchEngine(DeviceIndex(slice_host_index(host, p0, p1)))
        p, s, c = eng.search_batch(Q, **kw)
        eng.check_flags()
        assert int(p[p >= 0].min()) >= p0 and int(p.max()) < p1            # global pids of this shard only
        lists.append((p, s, c))
    gp = torch.stack([l[0] for l in lists]); gs = torch.stack([l[1] for l in lists]); gc = torch.stack([l[2] for l in lists])
    mp, ms, mc = sharded.merge_topk(gs, gp, gc, k)
    for b in range(Q.shape[0]):
        allp = torch.cat([gp[r, b, :gc[r, b]] for r in range(G)]).cpu()
        alls = torch.cat([gs[r, b, :gc[r, b]] for r in range(G)]).cpu()
        rp, rs = po.select_top(allp, alls, k)
        assert torch.equal(mp[b, :int(mc[b])].cpu(), rp) and torch.equal(ms[b, :int(mc[b])].cpu(), rs)
        assert int(mp[b, 0]) == int(g[f"rank_pids_{b}"][0])                # the planted passage still leads


def test_exhaustive_search_matches_padded_scores(pkg):
    """The executor's index-free branch (FLMR_base_executor.py:918-990): all queries x all items, best K."""
    g = torch.Generator().manual_seed(11)
    nQ, n_items, Ld, Lq, K = 5, 37, 50, 64, 10
    Q = torch.nn.functional.normalize(torch.randn(nQ, Lq, 128, generator=g), dim=-1)
    D = torch.nn.functional.normalize(torch.randn(n_items, Ld, 128, generator=g), dim=-1)
    lens = torch.randint(5, Ld + 1, (n_items,), generator=g)
    mask = (torch.arange(Ld).unsqueeze(0) < lens.unsqueeze(1)).unsqueeze(-1)
    out = pkg.exhaustive_search(Q, D, mask, K, truncate_scores=False)
    assert sorted(out.keys()) == list(range(nQ))
    for qi in range(nQ):
        ref = po.colbert_score(Q[qi:qi + 1].bfloat16().float().expand(n_items, -1, -1), D.bfloat16().float(), mask.squeeze(-1))
        rp, rs = po.select_top(torch.arange(n_items, dtype=torch.int32), ref, K)
        got = out[qi]
        assert [r for _, r, _ in got] == list(range(K))
        gs = torch.tensor([s for _, _, s in got])
        torch.testing.assert_close(gs, rs, rtol=2e-5, atol=2e-4)
        # same items unless two reference scores are closer than the summation-order noise
        assert [i for i, _, _ in got] == rp.tolist() or (rs[:-1] - rs[1:]).min() < 1e-4
    trunc = pkg.exhaustive_search(Q[:1], D, mask, 3)
    assert all(isinstance(s, int) for _, _, s in trunc[0])          # the reference stores int(score) here


def test_host_queries_are_fed_chunk_by_chunk(pkg, golden):
    """search_batch on HOST query embeddings (pinned or not): the copy-stream feed over several chunks returns
    exactly what the device-resident call returns."""
    from reranking_multimodal_retrievers_b200 import synthetic
    from reranking_multimodal_retrievers_b200.engine import SearchEngine
    from reranking_multimodal_retrievers_b200.index import DeviceIndex
    sx = synthetic.make_synthetic_index(1500, 10, 50, 2, seed=21, num_centroids=512, mode="codes", device="cuda")
    Q = synthetic.make_queries(sx, 37, 64, seed=22)
    eng = SearchEngine(DeviceIndex(sx), max_chunk=8)              # 5 chunks, the last one ragged
    want = eng.search_batch(Q.cuda(), k=20, ndocs=128)
    for host in (Q.cpu(), Q.cpu().pin_memory(), Q.cpu().double()):
        got = eng.search_batch(host, k=20, ndocs=128)
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(got, want))
    eng.check_flags()


def _oracle_score_pids(ix, q, cand, S, thr, ndocs):
    """IndexScorer.score_pids (index_storage.py:100-184) on an explicit candidate list."""
    idx = po.centroid_mask(S, thr)
    p2 = po.filter_pids(ix, cand, S, idx, ndocs)
    D = po.normalize(po.decompress_residuals(ix, p2))
    return p2, po.colbert_score_packed(q, D, ix.doclens[p2.long()])


def test_rank_with_filter_fn_more_candidates_than_ndocs(pkg):
    """IndexScorer.rank / score_pids with a filter_fn and with a caller-made pid list, n_candidates > ndocs (the
    normal case on real indexes): stage-1 pruning must act on the FILTERED list.  Against the oracle fed the
    centroid-score table `retrieve` returned."""
    from reranking_multimodal_retrievers_b200 import synthetic
    from reranking_multimodal_retrievers_b200.infra import ColBERTConfig
    sx = synthetic.make_synthetic_index(3000, 8, 40, 2, seed=31, num_centroids=256, mode="codes")
    Q = synthetic.make_queries(sx, 3, 48, seed=32)
    ix = po.OracleIndex(centroids=sx.centroids, bucket_weights=sx.bucket_weights, codes=sx.codes, residuals=sx.residuals,
                        doclens=sx.doclens, ivf=sx.ivf, ivf_lengths=sx.ivf_lengths, nbits=2)
    scorer = pkg.IndexScorer(sx)
    ndocs, thr = 64, 0.4
    cfg = ColBERTConfig(ncells=2, centroid_score_threshold=thr, ndocs=ndocs)
    keep_fn = lambda pids: pids[(pids % 3) != 1]                                   # noqa: E731
    for b in range(Q.shape[0]):
        q = Q[b]
        cand, S = scorer.retrieve(cfg, q.unsqueeze(0))
        cand2, S2 = scorer.retrieve(cfg, Q[(b + 1) % 3].unsqueeze(0))               # a later retrieve must not disturb ...
        Sc = S.float().cpu().contiguous()
        assert cand.numel() > 4 * ndocs                                             # ... the first one's tensors
        assert torch.equal(cand.cpu(), po.candidate_pids(ix, po.get_cells(Sc, 2)))
        for pid_list in (keep_fn(cand), cand, cand.flip(0)[: 3 * ndocs].contiguous()):
            scores, pids = scorer.score_pids(cfg, q.unsqueeze(0), pid_list, S)
            rp, rs = _oracle_score_pids(ix, q, pid_list.cpu(), Sc, thr, ndocs)
            assert torch.equal(pids.cpu(), rp), "stage-2 pids of a filtered candidate list differ from the oracle"
            assert ((scores.cpu() - rs).abs() <= SCORE_REL_TOL * rs.abs() + 1e-5).all()
        p, s = scorer.rank(cfg, q.unsqueeze(0), filter_fn=keep_fn)
        assert all(pid % 3 != 1 for pid in p) and s == sorted(s, reverse=True) and len(p) == ndocs // 4


@pytest.mark.parametrize("qml", [16, 5])
def test_query_maxlen_below_32(pkg, qml):
    """An index / config with query_maxlen < 32: only the first query_maxlen tokens drive candidate generation and
    the filter (`Q[:, :config.query_maxlen]`, index_storage.py:77), all tokens the exact MaxSim."""
    from reranking_multimodal_retrievers_b200 import synthetic
    from reranking_multimodal_retrievers_b200.engine import SearchEngine
    from reranking_multimodal_retrievers_b200.index import DeviceIndex
    sx = synthetic.make_synthetic_index(1500, 10, 60, 2, seed=41, num_centroids=512, mode="codes")
    Q = synthetic.make_queries(sx, 5, 64, seed=42)
    ix = po.OracleIndex(centroids=sx.centroids, bucket_weights=sx.bucket_weights, codes=sx.codes, residuals=sx.residuals,
                        doclens=sx.doclens, ivf=sx.ivf, ivf_lengths=sx.ivf_lengths, nbits=2)
    k, ndocs, ncells, thr = 20, 128, 2, 0.45
    eng = SearchEngine(DeviceIndex(sx), query_maxlen=qml)
    pids, scores, counts = eng.search_batch(Q, k=k, ncells=ncells, centroid_score_threshold=thr, ndocs=ndocs, keep_taps=True)
    eng.check_flags()
    t = eng.last_taps
    full = SearchEngine(DeviceIndex(sx)).search_batch(Q, k=k, ncells=ncells, centroid_score_threshold=thr, ndocs=ndocs)
    assert not torch.equal(full[0], pids)                      # the knob changes the outcome on this data
    for b in range(Q.shape[0]):
        S = t.S[b, :, :qml].float().cpu().contiguous()
        r = po.rank(ix, Q[b], ncells, thr, ndocs, query_maxlen=qml, S_override=S, taps=True)
        nc, n1, n2 = int(t.cand_counts[b]), int(t.stage1_counts[b]), int(t.stage2_counts[b])
        assert torch.equal(t.cand_pids[b, :nc].cpu(), r["candidates"])
        assert torch.equal(t.stage1_pids[b, :n1].cpu(), r["stage1_pids"])
        assert torch.equal(t.stage1_scores[b, :n1].cpu(), r["stage1_scores"])
        assert torch.equal(t.stage2_pids[b, :n2].cpu(), r["stage2_pids"])
        sc = t.scores[b, :n2].cpu()
        assert ((sc - r["scores_unsorted"]).abs() <= SCORE_REL_TOL * r["scores_unsorted"].abs() + 1e-5).all()
    with pytest.raises(pkg.PlaidError):
        SearchEngine(DeviceIndex(sx), query_maxlen=64)


def test_reference_format_index_streams_into_device_buffers(pkg, tmp_path):
    """load_reference_index(device=cuda): chunk files read by worker threads and copied through pinned staging buffers
    straight into the final device tensors (whole index and a pid-range shard) -- same bytes as the CPU loader."""
    from reranking_multimodal_retrievers_b200 import synthetic
    from reranking_multimodal_retrievers_b200.index import DeviceIndex, load_reference_index
    sx = synthetic.make_synthetic_index(2500, 5, 60, 4, seed=51, num_centroids=256, mode="codes")
    path = synthetic.write_reference_format(sx, str(tmp_path / "idx.nbits=4"), chunk_passages=333)   # 8 chunk files, ragged last
    for rng in (None, (400, 1999)):
        cpu = load_reference_index(path, rng)
        dev = load_reference_index(path, rng, device="cuda", workers=3)
        assert dev.codes.is_cuda and dev.residual_storage is not None
        assert torch.equal(dev.codes.cpu(), cpu.codes) and torch.equal(dev.residuals.cpu(), cpu.residuals)
        assert torch.equal(dev.doclens, cpu.doclens) and dev.pid_base == cpu.pid_base
        ix = DeviceIndex(dev)
        assert ix._res_storage.data_ptr() == dev.residual_storage.data_ptr()          # adopted, not copied
        assert torch.count_nonzero(ix._res_storage[-512:]) == 0
