"""BASELINE-sized runs (cfg1, cfg2 shapes of SURVEY 8d): size-independent properties on the whole batch
plus an oracle comparison on a few queries (the oracle needs ~50 ms per query at these sizes)."""
import pytest
import torch

from oracle import plaid_oracle as po

pytestmark = pytest.mark.gpu


def _build(N, B, Lq, nbits=2, lo=120, hi=239, seed=1234):
    from reranking_multimodal_retrievers_b200 import synthetic
    from reranking_multimodal_retrievers_b200.engine import SearchEngine
    from reranking_multimodal_retrievers_b200.index import DeviceIndex
    sx = synthetic.make_synthetic_index(N, lo, hi, nbits, seed=seed, mode="codes", device="cuda")
    Q, gold = synthetic.make_queries(sx, B, Lq, seed=99, return_gold=True)
    return sx, Q, gold, SearchEngine(DeviceIndex(sx))


def _oracle_index(sx):
    c = sx.cpu()
    return po.OracleIndex(centroids=c.centroids, bucket_weights=c.bucket_weights, codes=c.codes, residuals=c.residuals,
                          doclens=c.doclens, ivf=c.ivf, ivf_lengths=c.ivf_lengths, nbits=c.nbits)


@pytest.mark.parametrize("N,B", [(10_000, 256), (112_000, 192)])
def test_fullsize_properties_and_oracle_sample(N, B):
    k = 100
    sx, Q, gold, eng = _build(N, B, 64)
    p1, s1, c1 = eng.search_batch(Q, k=k, keep_taps=True)
    eng.check_flags()
    t = eng.last_taps
    # every query returns k results; scores descending, ties by pid descending
    assert torch.all(c1 == k)
    assert torch.all(s1[:, :-1] >= s1[:, 1:])
    tie = s1[:, :-1] == s1[:, 1:]
    assert torch.all(p1[:, :-1][tie] > p1[:, 1:][tie])
    # no passage twice in a list, all pids valid
    srt = p1.sort(dim=1).values
    assert torch.all(srt[:, 1:] != srt[:, :-1]) and int(p1.min()) >= 0 and int(p1.max()) < N
    # the planted passage leads
    assert float((p1[:, 0].cpu() == gold.cpu().to(torch.int32)).float().mean()) >= 0.99
    # idempotence, and independence from how the batch is chunked
    p2, s2, c2 = eng.search_batch(Q, k=k)
    assert torch.equal(p1, p2) and torch.equal(s1, s2)
    eng.max_chunk = 64
    eng._ws_key = None
    p3, s3, c3 = eng.search_batch(Q, k=k)
    assert torch.equal(p1, p3) and torch.equal(s1, s3)
    # stage containment on the last chunk: stage 2 within stage 1 within the candidates
    for b in range(0, 64, 16):
        cand = set(t.cand_pids[b, :int(t.cand_counts[b])].tolist())
        st1 = set(t.stage1_pids[b, :int(t.stage1_counts[b])].tolist())
        st2 = set(t.stage2_pids[b, :int(t.stage2_counts[b])].tolist())
        assert st2 <= st1 <= cand and len(st1) == 1024 and len(st2) == 256
    # oracle on a sample, fed our centroid-score table: integer stages bit-exact, scores within 1e-3
    ix = _oracle_index(sx)
    eng.max_chunk = 512
    eng._ws_key = None
    pids, scores, counts = eng.search_batch(Q[:8], k=k, keep_taps=True)
    t = eng.last_taps
    Qc = Q.cpu()
    for b in (0, 3, 7):
        S = t.S[b].float().cpu().contiguous()
        r = po.rank(ix, Qc[b], 2, 0.45, 1024, S_override=S, taps=True)
        assert torch.equal(t.cand_pids[b, :int(t.cand_counts[b])].cpu(), r["candidates"])
        assert torch.equal(t.stage1_pids[b, :1024].cpu(), r["stage1_pids"])
        assert torch.equal(t.stage2_pids[b, :256].cpu(), r["stage2_pids"])
        sc = t.scores[b, :256].cpu()
        assert ((sc - r["scores_unsorted"]).abs() <= 1e-3 * r["scores_unsorted"].abs() + 1e-6).all()
        rp, rs = po.select_top(r["stage2_pids"], sc, k)
        assert torch.equal(pids[b].cpu(), rp) and torch.equal(scores[b].cpu(), rs)


def test_resident_batch_runs_as_one_chunk_and_matches_the_host_fed_chunks():
    """Device-resident batches of up to 1024 queries are searched as ONE chunk; host-fed ones cross PCIe in equal pieces
    (query prep + centroid scoring per piece, the rest once) or, with host_piece = 0, as whole chunks one ahead of the
    search: all three paths return the same lists bit for bit."""
    k, B = 100, 640
    sx, Q, gold, eng = _build(20_000, B, 64)
    assert eng.chunk_size(B, resident=True) == B and eng.chunk_size(B) == 320     # whole-chunk feeding: two equal chunks
    pd, sd, cd = eng.search_batch(Q, k=k)                       # one chunk of 640
    eng.check_flags()
    Qh = Q.cpu().pin_memory()
    ph, sh, ch = eng.search_batch(Qh, k=k)                      # 3 pieces of 216 / 216 / 208 queries, one chunk
    eng.check_flags()
    assert torch.equal(pd, ph) and torch.equal(sd, sh) and torch.equal(cd, ch)
    eng.host_piece = 0
    pw, sw, cw = eng.search_batch(Qh, k=k)                      # 320 + 320, H2D one chunk ahead
    eng.check_flags()
    assert torch.equal(pd, pw) and torch.equal(sd, sw) and torch.equal(cd, cw)
    assert float((pd[:, 0].cpu() == gold.cpu().to(torch.int32)).float().mean()) >= 0.99


def test_colbert_score_host_query_batch_is_copied_in_pieces_same_scores():
    """A large HOST batch of query matrices (the rerank hand-off) crosses PCIe in ~16 MB pieces behind the MaxSim of the
    previous piece: the scores equal those of the device-resident call bit for bit (ragged last piece, n < nQ * dpq)."""
    import reranking_multimodal_retrievers_b200 as pkg
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    nQ, dpq, Ld, Lq = 1100, 2, 16, 64
    n = nQ * dpq - 1
    Q = torch.nn.functional.normalize(torch.randn(nQ, Lq, 128, generator=g, device="cuda"), dim=-1)
    D = torch.nn.functional.normalize(torch.randn(n, Ld, 128, generator=g, device="cuda"), dim=-1).bfloat16()
    lens = torch.randint(1, Ld + 1, (n,), generator=g, device="cuda")
    mask = torch.arange(Ld, device="cuda").unsqueeze(0) < lens.unsqueeze(1)
    dev_scores = pkg.colbert_score(Q, D, mask, docs_per_query=dpq)
    host_scores = pkg.colbert_score(Q.cpu().pin_memory(), D, mask, docs_per_query=dpq)
    assert host_scores.is_cuda and torch.equal(dev_scores, host_scores)


def test_fullsize_padded_rerank_linearity_and_permutation():
    """cfg5 shape (top-100 rerank, bf16 passage embeddings): properties of the padded MaxSim on a
    4096-passage slab -- permutation equivariance over passages, invariance to padding content,
    and agreement with the packed form on the same tokens."""
    import reranking_multimodal_retrievers_b200 as pkg
    g = torch.Generator().manual_seed(5)
    nQ, dpq, Ld, Lq = 41, 100, 240, 64
    n = nQ * dpq - 7                                    # last query has a short list
    Q = torch.nn.functional.normalize(torch.randn(nQ, Lq, 128, generator=g), dim=-1).cuda()
    D = torch.nn.functional.normalize(torch.randn(n, Ld, 128, generator=g), dim=-1).cuda().bfloat16()
    lens = torch.randint(120, Ld + 1, (n,), generator=g).cuda()
    mask = torch.arange(Ld, device="cuda").unsqueeze(0) < lens.unsqueeze(1)
    s = pkg.colbert_score(Q, D, mask, docs_per_query=dpq)
    # padding content is irrelevant (masked positions count as -9999)
    D2 = D.clone()
    D2[~mask] = 7.0
    assert torch.equal(pkg.colbert_score(Q, D2, mask, docs_per_query=dpq), s)
    # permuting the passages of one query permutes its scores
    perm = torch.randperm(dpq, generator=g).cuda()
    D3, m3 = D.clone(), mask.clone()
    D3[:dpq], m3[:dpq] = D[:dpq][perm], mask[:dpq][perm]
    s3 = pkg.colbert_score(Q, D3, m3, docs_per_query=dpq)
    assert torch.equal(s3[:dpq], s[:dpq][perm]) and torch.equal(s3[dpq:], s[dpq:])
    # the packed form on the same (unpadded) tokens: same maxima unless a maximum is negative (clamp at 0)
    q0 = Q[:1]
    packed = torch.cat([D[i, :int(lens[i])] for i in range(dpq)])
    sp = pkg.colbert_score_packed(q0, packed, lens[:dpq])
    full = (D[:dpq].float() @ q0[0].bfloat16().float().T)
    full[~mask[:dpq]] = -9999
    ref_clamped = full.max(1).values.clamp(min=0).sum(-1)
    torch.testing.assert_close(sp, ref_clamped, rtol=2e-5, atol=2e-4)
    torch.testing.assert_close(s[:dpq], full.max(1).values.sum(-1), rtol=2e-5, atol=2e-4)


def _oracle_check(eng, sx, Q, k, queries, ndocs=1024):
    """Oracle fed OUR centroid-score table on a few queries: integer stages bit-exact, scores within 1e-3, final
    order exactly (score desc, pid desc) of our scores."""
    ix = _oracle_index(sx)
    pids, scores, counts = eng.search_batch(Q[:max(queries) + 1], k=k, keep_taps=True)
    eng.check_flags()
    t = eng.last_taps
    Qc = Q.cpu()
    for b in queries:
        S = t.S[b].float().cpu().contiguous()
        r = po.rank(ix, Qc[b], 2, 0.45, ndocs, S_override=S, taps=True)
        assert torch.equal(t.cand_pids[b, :int(t.cand_counts[b])].cpu(), r["candidates"])
        n1, n2 = int(t.stage1_counts[b]), int(t.stage2_counts[b])
        assert torch.equal(t.stage1_pids[b, :n1].cpu(), r["stage1_pids"])
        assert torch.equal(t.stage1_scores[b, :n1].cpu(), r["stage1_scores"])
        assert torch.equal(t.stage2_pids[b, :n2].cpu(), r["stage2_pids"])
        assert torch.equal(t.stage2_scores[b, :n2].cpu(), r["stage2_scores"])
        sc = t.scores[b, :n2].cpu()
        assert ((sc - r["scores_unsorted"]).abs() <= 1e-3 * r["scores_unsorted"].abs() + 1e-6).all()
        rp, rs = po.select_top(r["stage2_pids"], sc, k)
        assert torch.equal(pids[b, :int(counts[b])].cpu(), rp) and torch.equal(scores[b, :int(counts[b])].cpu(), rs)


def test_cfg3_at_size_long_queries_nbits4():
    """BASELINE configs[2] at size: 100k passages of 128..512 tokens (32M tokens, C = 65536), 4-bit residuals, 320-token
    PreFLMR queries (three query m-tiles in the fused MaxSim; candidate stage on the first 32 tokens)."""
    k = 100
    sx, Q, gold, eng = _build(100_000, 48, 320, nbits=4, lo=128, hi=512, seed=1237)
    p1, s1, c1 = eng.search_batch(Q, k=k)
    eng.check_flags()
    assert torch.all(c1 == k) and torch.all(s1[:, :-1] >= s1[:, 1:])
    srt = p1.sort(dim=1).values
    assert torch.all(srt[:, 1:] != srt[:, :-1]) and int(p1.min()) >= 0 and int(p1.max()) < 100_000
    assert float((p1[:, 0].cpu() == gold.cpu().to(torch.int32)).float().mean()) >= 0.99
    eng.max_chunk = 16                                  # chunking independence
    p2, s2, _ = eng.search_batch(Q, k=k)
    assert torch.equal(p1, p2) and torch.equal(s1, s2)
    eng.max_chunk = 512
    unf = type(eng)(eng.index, fused=False)             # the unfused pair builds the same tiles: same bits
    p3, s3, _ = unf.search_batch(Q[:16], k=k)
    assert torch.equal(p1[:16], p3) and torch.equal(s1[:16], s3)
    _oracle_check(eng, sx, Q, k, (0, 5, 11))


def test_cfg4_shard_at_size_large_codebook():
    """One 1/8 shard of BASELINE configs[3]: 1.25M passages (225M tokens) against the full collection's C = 524288
    centroids -- the large-codebook chunking of the engine (148-query chunks, several centroid ranges per group) and
    33 MB score tables per query."""
    from reranking_multimodal_retrievers_b200 import synthetic
    from reranking_multimodal_retrievers_b200.engine import SearchEngine
    from reranking_multimodal_retrievers_b200.index import DeviceIndex
    k, N, C = 100, 1_250_000, 524_288
    sx = synthetic.make_synthetic_index(N, 120, 239, 2, seed=1238, mode="codes", device="cuda", num_centroids=C)
    Q, gold = synthetic.make_queries(sx, 300, 64, seed=99, return_gold=True)
    eng = SearchEngine(DeviceIndex(sx), s_budget_bytes=6 << 30)
    assert eng.chunk_size(300) == 148                    # 148-query chunks (4 centroid ranges per query group), the last one ragged
    p1, s1, c1 = eng.search_batch(Q, k=k)
    eng.check_flags()
    assert torch.all(c1 == k) and torch.all(s1[:, :-1] >= s1[:, 1:])
    srt = p1.sort(dim=1).values
    assert torch.all(srt[:, 1:] != srt[:, :-1]) and int(p1.min()) >= 0 and int(p1.max()) < N
    assert float((p1[:, 0].cpu() == gold.cpu().to(torch.int32)).float().mean()) >= 0.99
    big = SearchEngine(eng.index)                        # default table budget: one chunk of 300 queries
    assert big.chunk_size(300) == 300
    pb, sb, _ = big.search_batch(Q, k=k)
    assert torch.equal(pb, p1) and torch.equal(sb, s1)
    del big
    sub = eng.search_batch(Q[37:61], k=k)                # a different chunking of the same queries
    assert torch.equal(sub[0], p1[37:61]) and torch.equal(sub[1], s1[37:61])
    _oracle_check(eng, sx, Q, k, (0, 2))
