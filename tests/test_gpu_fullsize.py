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
