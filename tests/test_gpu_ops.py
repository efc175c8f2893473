"""GPU parity of the reference-named operators, called through the C ABI, against the CPU oracle and
the golden vectors recorded from the reference.  Integer / byte / index results: bit-exact."""
import numpy as np
import pytest
import torch

from plaid_test_helpers import golden_oracle_index, nonzero_rows
from oracle import plaid_oracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import reranking_multimodal_retrievers_b200 as pkg
    from reranking_multimodal_retrievers_b200 import ops
    assert torch.cuda.is_available()
    return pkg, ops


def test_prepare_queries_zero_rows_and_bf16(P):
    _, ops = P
    g = torch.Generator().manual_seed(3)
    Q = torch.randn(5, 40, 128, generator=g)
    Q[0, 3] = 0
    Q[0, 10:14] = 0
    Q[2, :] = 0          # a query that is entirely zero
    Q[4, 39] = 0
    Qb, qlens = ops.prepare_queries(Q, remove_zero_rows=True)
    assert Qb.shape == (8, 64, 128) and qlens.shape == (8,)
    for b in range(5):
        kept = nonzero_rows(Q[b])
        assert int(qlens[b]) == kept.shape[0]
        assert torch.equal(Qb[b, :kept.shape[0]].cpu(), kept.bfloat16())       # RNE rounding, order preserved
        assert torch.count_nonzero(Qb[b, kept.shape[0]:]) == 0
    assert qlens[5:].tolist() == [0, 0, 0] and torch.count_nonzero(Qb[5:]) == 0
    Qb2, qlens2, Qh2 = ops.prepare_queries(Q, remove_zero_rows=False, with_f16=True)
    assert qlens2[:5].tolist() == [40] * 5
    assert torch.equal(Qb2[:5, :40].cpu(), Q.bfloat16())
    assert torch.equal(Qh2[:5, :40].cpu(), Q.half()) and torch.count_nonzero(Qh2[5:]) == 0      # fp16 twin, same rows
    assert torch.count_nonzero(Qh2[:5, 40:]) == 0


def test_decompress_residuals_bit_exact_vs_golden(P, golden):
    _, ops = P
    g = golden
    ix = golden_oracle_index(g)
    pids = torch.from_numpy(g["stage2_0"])
    out = ops.decompress_residuals(pids, ix.doclens, ix.offsets, ix.bucket_weights, ix.reversed_bit_map, ix.lookup,
                                   ix.residuals, ix.codes, ix.centroids, 128, ix.nbits)
    assert out.dtype == torch.float32
    assert np.array_equal(out.cpu().numpy(), g["D_0"])                       # recorded from the reference
    assert torch.equal(out.cpu(), po.decompress_residuals(ix, pids))


@pytest.mark.parametrize("nbits", [1, 2, 4, 8])
def test_decompress_and_unpack_all_bit_widths(P, nbits):
    _, ops = P
    from reranking_multimodal_retrievers_b200.synthetic import make_synthetic_index
    sx = make_synthetic_index(300, 1, 70, nbits, seed=40 + nbits, num_centroids=256, mode="codes")
    ix = po.OracleIndex(centroids=sx.centroids, bucket_weights=sx.bucket_weights, codes=sx.codes,
                        residuals=sx.residuals, doclens=sx.doclens, ivf=sx.ivf, ivf_lengths=sx.ivf_lengths, nbits=nbits)
    pids = torch.tensor([0, 299, 17, 17, 5, 123, 64], dtype=torch.int32)     # unordered, with a repeat
    out = ops.decompress_residuals(pids, ix.doclens, ix.offsets, ix.bucket_weights, ix.reversed_bit_map, ix.lookup,
                                   ix.residuals, ix.codes, ix.centroids, 128, nbits)
    assert torch.equal(out.cpu(), po.decompress_residuals(ix, pids))
    codes = ops.unpack_residual_codes(ix.residuals[:500], nbits, ix.reversed_bit_map, ix.lookup)
    assert torch.equal(codes.cpu(), po.unpack_residual_codes(ix, ix.residuals[:500]))   # integer: bit-exact
    empty = ops.decompress_residuals(pids[:0], ix.doclens, ix.offsets, ix.bucket_weights, ix.reversed_bit_map,
                                     ix.lookup, ix.residuals, ix.codes, ix.centroids, 128, nbits)
    assert empty.shape == (0, 128)


def test_filter_pids_bit_exact_with_reference_S(P, golden):
    """Injected-input protocol: the reference's own centroid-score table goes in, the
    stage-2 pid list that comes out must equal the reference's, order included."""
    _, ops = P
    g = golden
    ix = golden_oracle_index(g)
    ndocs = int(g["ndocs"])
    for b in (0, 1):
        S, idx, cand = (torch.from_numpy(g[f"{n}_{b}"]) for n in ("S", "idx", "cand"))
        p2, (p1, s1, s2) = ops.filter_pids(cand, S, ix.codes, ix.doclens, ix.offsets, idx, ndocs, return_stages=True)
        assert np.array_equal(p2.cpu().numpy(), g[f"stage2_{b}"])
        o2, (o1, os1, os2) = po.filter_pids(ix, cand, S, idx, ndocs, return_stages=True)
        assert torch.equal(p1.cpu(), o1) and torch.equal(p2.cpu(), o2)
        assert torch.equal(s1.cpu(), os1) and torch.equal(s2.cpu(), os2)     # sequential fp32 sums: bit-exact
        a = ops.approx_scores(cand, S, ix.codes, ix.offsets, idx)
        assert torch.equal(a.cpu(), po.approx_scores(ix, cand, S, idx))
        a = ops.approx_scores(cand, S, ix.codes, ix.offsets, None)
        assert torch.equal(a.cpu(), po.approx_scores(ix, cand, S, None))


def test_filter_pids_fewer_candidates_than_ndocs_and_short_queries(P, golden):
    _, ops = P
    g = golden
    ix = golden_oracle_index(g)
    S, idx = torch.from_numpy(g["S_0"]), torch.from_numpy(g["idx_0"])
    cand = torch.from_numpy(g["cand_0"])[:50]
    p2, (p1, s1, s2) = ops.filter_pids(cand, S, ix.codes, ix.doclens, ix.offsets, idx, 128, return_stages=True)
    o2, (o1, os1, os2) = po.filter_pids(ix, cand, S, idx, 128, return_stages=True)
    assert p1.numel() == 50 and p2.numel() == 32
    assert torch.equal(p1.cpu(), o1) and torch.equal(p2.cpu(), o2) and torch.equal(s2.cpu(), os2)
    # nq < 32 query tokens (PreFLMR masks instruction tokens): the sum runs over nq columns only
    S7 = S[:, :7].contiguous()
    idx7 = S7.max(-1).values >= 0.45
    cand = torch.from_numpy(g["cand_0"])
    p2 = ops.filter_pids(cand, S7, ix.codes, ix.doclens, ix.offsets, idx7, 64)
    assert torch.equal(p2.cpu(), po.filter_pids(ix, cand, S7, idx7, 64))
    # no candidates at all
    p2 = ops.filter_pids(cand[:0], S, ix.codes, ix.doclens, ix.offsets, idx, 64)
    assert p2.numel() == 0


@pytest.mark.parametrize("nq", [32, 7, 1])
def test_approx_scores_fp16_table(P, nq):
    """The fp16 score table (the engine's default, candidate_generation.py:52): stage 2 reads two rows per warp load and
    keeps half2 maxima -- maxima of fp16 values are exact, so the scores equal the oracle's on the rounded table bit for
    bit.  Passage lengths cover empty, < 32, multiples of 32 and odd tails."""
    _, ops = P
    gen = torch.Generator().manual_seed(7)
    C = 384
    doclens = torch.tensor([0, 1, 2, 31, 32, 33, 63, 64, 65, 96, 127, 180, 5, 0, 40, 17] * 3, dtype=torch.int64)
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64), doclens.cumsum(0)])
    codes = torch.randint(0, C, (int(offsets[-1]),), generator=gen, dtype=torch.int32)
    ix = po.OracleIndex(centroids=torch.zeros(C, 128), bucket_weights=torch.zeros(4), codes=codes,
                        residuals=torch.zeros(codes.numel(), 32, dtype=torch.uint8), doclens=doclens,
                        ivf=torch.zeros(1, dtype=torch.int32), ivf_lengths=torch.zeros(C, dtype=torch.int64), nbits=2)
    S = torch.randn(C, nq, generator=gen)
    Sr = S.half().float()
    pids = torch.randperm(doclens.numel(), generator=gen).to(torch.int32)
    a = ops.approx_scores(pids, S, ix.codes, ix.offsets, None, table_f16=True)
    assert torch.equal(a.cpu(), po.approx_scores(ix, pids, Sr, None))
    idx = Sr.max(-1).values >= 1.2                     # a sparse pruning mask: many tokens without a surviving centroid
    a = ops.approx_scores(pids, S, ix.codes, ix.offsets, idx, table_f16=True)
    assert torch.equal(a.cpu(), po.approx_scores(ix, pids, Sr, idx))


def test_select_top_order_and_ties(P):
    _, ops = P
    g = torch.Generator().manual_seed(9)
    n = 5000
    scores = torch.randint(0, 40, (n,), generator=g).float() * 0.25 - 3.0      # heavy ties, negatives
    scores[17] = float("-inf")
    pids = torch.randperm(1 << 20, generator=g)[:n].to(torch.int32)
    for keep in (1, 7, 256, 1024, 4096, 6000):
        op, os_ = ops.select_top(pids, scores, keep)
        rp, rs = po.select_top(pids, scores, keep)
        assert torch.equal(op.cpu(), rp) and torch.equal(os_.cpu(), rs)
    # longer lists: the bucket of the keep-th key is collected into shared memory, narrowed there by further radix
    # passes, squeezed and finished by rank counting (tied scores push the decision into the pid bytes)
    n = 20000
    for levels, scale in ((40, 0.25), (3, 1.0), (20000, 1e-3)):
        scores = torch.randint(0, levels, (n,), generator=g).float() * scale - 3.0
        pids = torch.randperm(1 << 17, generator=g)[:n].to(torch.int32)
        for keep in (100, 1024, 4096, 8192, 16384):
            op, os_ = ops.select_top(pids, scores, keep)
            rp, rs = po.select_top(pids, scores, keep)
            assert torch.equal(op.cpu(), rp) and torch.equal(os_.cpu(), rs), (levels, keep)
    # lists of a multi-million-passage shard (>= 32768 keys: warp-aggregated histogram updates): most keys carry the
    # "no surviving centroid" score of stage 1, a few thousand something better
    n = 150000
    scores = torch.full((n,), -9999.0 * 32)
    live = torch.randperm(n, generator=g)[:6000]
    scores[live] = torch.randint(0, 500, (6000,), generator=g).float() * 0.125 - 20.0
    pids = torch.randperm(1 << 22, generator=g)[:n].to(torch.int32)
    for keep in (1024, 4096, 8192):
        op, os_ = ops.select_top(pids, scores, keep)
        rp, rs = po.select_top(pids, scores, keep)
        assert torch.equal(op.cpu(), rp) and torch.equal(os_.cpu(), rs), keep


def test_segmented_maxsim_and_lookup(P, golden):
    _, ops = P
    gen = torch.Generator().manual_seed(5)
    lengths = torch.tensor([3, 1, 170, 0, 9, 64], dtype=torch.long)
    for nq in (32, 40, 64, 128):
        scores = torch.randn(int(lengths.sum()), nq, generator=gen)
        out = ops.segmented_maxsim(scores, lengths)
        assert torch.equal(out.cpu(), po.segmented_maxsim(scores, lengths))  # same left-to-right fp32 sum
        assert float(out[3]) == 0.0
    ix = golden_oracle_index(golden)
    pids = torch.tensor([5, 0, 17, 3, 5], dtype=torch.long)
    lengths, offsets = ix.doclens[pids], ix.offsets[pids]
    for table in (ix.residuals, ix.codes):
        out = ops.segmented_lookup(table, pids, lengths, offsets)
        ref = torch.cat([table[o:o + l] for o, l in zip(offsets.tolist(), lengths.tolist())])
        assert torch.equal(out.cpu(), ref)


def test_colbert_score_reduce(P):
    pkg, _ = P
    g = torch.Generator().manual_seed(12)
    sp = torch.randn(7, 19, 45, generator=g)
    mask = torch.rand(7, 19, generator=g) > 0.3
    mask[2] = False                                  # fully padded passage -> -9999 * Lq
    out = pkg.colbert_score_reduce(sp, mask)
    ref, _ = po.colbert_score_reduce(sp, mask)
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-6, atol=1e-4)
    s, raw = pkg.flmr_colbert_score_reduce(sp, mask)
    assert torch.equal(raw.cpu(), po.colbert_score_reduce(sp, mask)[1])


def test_merge_topk_matches_cpu_merge(P):
    from reranking_multimodal_retrievers_b200 import sharded
    g = torch.Generator().manual_seed(21)
    G, B, k = 8, 33, 100
    scores = torch.randint(0, 500, (G, B, k), generator=g).float() / 7.0
    scores = scores.sort(dim=-1, descending=True).values
    pids = torch.randperm(G * B * k, generator=g).reshape(G, B, k).to(torch.int32)
    counts = torch.randint(0, k + 1, (G, B), generator=g).to(torch.int32)
    counts[0, 0] = k
    counts[:, 1] = 0                                    # a query for which no shard found anything
    op, os_, oc = sharded.merge_topk(scores.cuda(), pids.cuda(), counts.cuda(), k)
    for b in range(B):
        allp = torch.cat([pids[gk, b, :counts[gk, b]] for gk in range(G)])
        alls = torch.cat([scores[gk, b, :counts[gk, b]] for gk in range(G)])
        rp, rs = po.select_top(allp, alls, k) if allp.numel() else (allp, alls)
        m = int(oc[b])
        assert m == rp.numel()
        assert torch.equal(op[b, :m].cpu(), rp) and torch.equal(os_[b, :m].cpu(), rs)
        assert torch.all(op[b, m:] == -1)
    # message packing used for the single all-gather
    msg = sharded.pack_lists(pids[0].cuda(), scores[0].cuda(), counts[0].cuda())
    p, s, c = sharded.unpack_lists(msg.unsqueeze(0), k)
    assert torch.equal(p[0].cpu(), pids[0]) and torch.equal(s[0].cpu(), scores[0]) and torch.equal(c[0].cpu(), counts[0])


def test_strided_tensor_lookup_and_padding(P, golden):
    """StridedTensor (strided_tensor.py:77-99, strided_tensor_core.py:85-96): the IVF and the code stream."""
    pkg, _ = P
    ix = golden_oracle_index(golden)
    ivf = pkg.StridedTensor(ix.ivf, ix.ivf_lengths)
    cells = torch.tensor([7, 0, 1023, 7, 512])
    pids, lens = ivf.lookup(cells)
    ref = torch.cat([ix.ivf[ix.ivf_offsets[c]:ix.ivf_offsets[c + 1]] for c in cells.tolist()])
    assert torch.equal(pids.cpu(), ref) and torch.equal(lens.cpu(), ix.ivf_lengths[cells])
    codes = pkg.StridedTensor(ix.codes, ix.doclens)
    docs = torch.tensor([3, ix.doclens.numel() - 1, 0])
    padded, mask = codes.lookup(docs, output="padded")
    assert padded.shape == (3, int(ix.doclens[docs].max())) and mask.shape == padded.shape
    for i, d in enumerate(docs.tolist()):
        L = int(ix.doclens[d])
        assert torch.equal(padded[i, :L].cpu(), ix.codes[ix.offsets[d]:ix.offsets[d] + L])
        assert not bool(mask[i, L:].any()) and bool(mask[i, :L].all()) and int(padded[i, L:].abs().sum()) == 0
    res = pkg.StridedTensor(ix.residuals, ix.doclens)
    full, fmask = res.as_padded_tensor()
    assert full.shape == (ix.doclens.numel(), int(ix.doclens.max()), ix.residuals.shape[1]) and fmask.dim() == 3
    assert torch.equal(full[5, :int(ix.doclens[5])].cpu(), ix.residuals[ix.offsets[5]:ix.offsets[6]])


@pytest.mark.parametrize("name", ["codec_nbits2", "codec_nbits4"])
def test_codec_compress_bit_exact_vs_reference_golden(P, name):
    """Index-build codec kernels (8f-3) against vectors recorded from the reference's ResidualCodec.compress."""
    import os
    import numpy as np
    from reranking_multimodal_retrievers_b200 import codec
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"{name}.npz"))
    nbits = int(g["nbits"])
    embs = torch.from_numpy(g["embs"]).cuda()
    cent = torch.from_numpy(g["centroids_f16"]).cuda()
    cut = torch.from_numpy(g["bucket_cutoffs"]).cuda()
    ref_codes = torch.from_numpy(g["codes"])
    ref_res = torch.from_numpy(g["residuals"])
    # residual / bucketize / pack given the reference's codes: bit-exact
    res = codec.compress_residuals(embs, ref_codes.cuda(), cent, cut, nbits)
    assert torch.equal(res.cpu(), ref_res)
    # argmax on the tensor cores: these embeddings sit next to their centroid, the codes are unambiguous
    codes, res2 = codec.compress(embs, cent, cut, nbits)
    assert torch.equal(codes.cpu(), ref_codes) and torch.equal(res2.cpu(), ref_res)


@pytest.mark.parametrize("nbits,n,C", [(1, 77, 64), (2, 1000, 1024), (8, 130, 32)])
def test_codec_all_bit_widths_and_argmax_vs_oracle(P, nbits, n, C):
    from reranking_multimodal_retrievers_b200 import codec
    g = torch.Generator().manual_seed(100 + nbits)
    cent = torch.nn.functional.normalize(torch.randn(C, 128, generator=g), dim=-1).half()
    embs = torch.nn.functional.normalize(torch.randn(n, 128, generator=g), dim=-1)       # far from any centroid: near-ties
    res_all = embs - cent.float()[torch.randint(0, C, (n,), generator=g)]
    if nbits < 8:
        cut = res_all.flatten().quantile(torch.linspace(0, 1, 2 ** nbits + 1)[1:-1])
    else:
        cut = torch.linspace(-0.3, 0.3, 255)
    codes = codec.compress_into_codes(embs.cuda(), cent.cuda()).cpu()
    # oracle on the same bf16-rounded operands; a different pick is only allowed where the two best are a near-tie
    S = cent.float().bfloat16().float() @ embs.bfloat16().float().T
    ref = S.max(dim=0).indices.to(torch.int32)
    diff = codes != ref
    if diff.any():
        top2 = S.topk(2, dim=0).values
        assert ((top2[0] - top2[1])[diff] < 1e-5).all()
        assert (S.gather(0, codes.long().unsqueeze(0))[0][diff] >= top2[1][diff] - 1e-6).all()
    _, want = po.codec_compress(cent, cut, nbits, embs, codes=codes)
    got = codec.compress_residuals(embs.cuda(), codes.cuda(), cent.cuda(), cut.cuda(), nbits)
    assert torch.equal(got.cpu(), want)


def test_codec_build_index_is_searchable(P):
    """Embeddings -> codes/residuals/IVF on the device -> the planted passage is found."""
    from reranking_multimodal_retrievers_b200 import codec
    from reranking_multimodal_retrievers_b200.engine import SearchEngine
    from reranking_multimodal_retrievers_b200.index import DeviceIndex
    g = torch.Generator().manual_seed(5)
    C, N, nbits = 256, 300, 2
    cent = torch.nn.functional.normalize(torch.randn(C, 128, generator=g), dim=-1).half()
    doclens = torch.randint(8, 40, (N,), generator=g)
    assign = torch.randint(0, C, (int(doclens.sum()),), generator=g)
    embs = torch.nn.functional.normalize(cent.float()[assign] + 0.05 * torch.randn(assign.numel(), 128, generator=g), dim=-1)
    res = embs - cent.float()[assign]
    cut = res.flatten().quantile(torch.tensor([0.25, 0.5, 0.75]))
    w = res.flatten().quantile(torch.tensor([0.125, 0.375, 0.625, 0.875]))
    hx = codec.build_index(embs.cuda(), doclens, cent.cuda(), cut.cuda(), w.cuda(), nbits)
    assert torch.equal(hx.codes.cpu(), assign.to(torch.int32))
    eng = SearchEngine(DeviceIndex(hx))
    off = torch.cat([torch.zeros(1, dtype=torch.int64), doclens.cumsum(0)])
    gold = torch.tensor([7, 123, 299])
    Q = torch.stack([torch.nn.functional.normalize(
        embs[off[p]: off[p] + 8].repeat(4, 1) + 0.02 * torch.randn(32, 128, generator=g), dim=-1) for p in gold.tolist()])
    pids, scores, counts = eng.search_batch(Q, k=5, ndocs=64)
    eng.check_flags()
    assert pids[:, 0].cpu().tolist() == gold.tolist()


@pytest.mark.parametrize("nbits", [1, 2, 4, 8])
def test_residual_codec_gpu_operators(P, nbits):
    """ResidualCodec.decompress_residuals (GPU form: token rows, fp16) and ResidualCodec.packbits
    (CB/indexing/codecs/residual.py:115,130): bit-exact against the oracle; decompress() within one half ulp."""
    pkg, ops = P
    from reranking_multimodal_retrievers_b200.synthetic import make_synthetic_index
    sx = make_synthetic_index(200, 1, 60, nbits, seed=60 + nbits, num_centroids=256, mode="codes")
    rbm, lut = po.codec_tables(nbits)
    n = sx.num_embeddings
    for m in (n, 1, 33):                                        # odd counts exercise the half-warp tail
        res, codes = sx.residuals[:m], sx.codes[:m]
        out = pkg.ResidualCodec.decompress_residuals(res, sx.bucket_weights.half(), rbm, lut, codes, sx.centroids, 128, nbits)
        want = po.codec_decompress_gpu_form(sx.bucket_weights, rbm, lut, res, codes, sx.centroids)
        assert out.dtype == torch.float16 and out.shape == (m, 128)
        assert torch.equal(out.cpu(), want)
    # one half rounding away from the fp32 CPU operator on the same bytes
    ix = po.OracleIndex(centroids=sx.centroids, bucket_weights=sx.bucket_weights.half().float(), codes=sx.codes,
                        residuals=sx.residuals, doclens=sx.doclens, ivf=sx.ivf, ivf_lengths=sx.ivf_lengths, nbits=nbits)
    D32 = po.decompress_residuals(ix, torch.arange(sx.num_passages, dtype=torch.int32))
    out = ops.codec_decompress_residuals(sx.residuals, sx.bucket_weights.half(), rbm, lut, sx.codes, sx.centroids, 128, nbits)
    assert torch.equal(out.cpu(), D32.half())
    # the codec object: decompress() = normalised half rows
    cfg = pkg.ColBERTConfig(dim=128, nbits=nbits)
    codec = pkg.ResidualCodec(cfg, sx.centroids, bucket_cutoffs=sx.bucket_cutoffs, bucket_weights=sx.bucket_weights)
    Dn = codec.decompress(pkg.ResidualEmbeddings(sx.codes, sx.residuals)).cpu()
    ref = po.codec_decompress_gpu_form(sx.bucket_weights, rbm, lut, sx.residuals, sx.codes, sx.centroids, normalize=True)
    assert (Dn.float() - ref.float()).abs().max() <= 2 ** -10 and (Dn != ref).float().mean() < 0.01
    # packbits / binarize
    g = torch.Generator().manual_seed(nbits)
    flags = torch.randint(0, 2, (8 * 1000,), generator=g, dtype=torch.uint8)
    assert torch.equal(pkg.ResidualCodec.packbits(flags).cpu(), po.codec_packbits(flags))
    flags[::7] *= 200                                           # any non-zero byte is a set flag
    assert torch.equal(ops.packbits(flags).cpu(), po.codec_packbits(flags))
    assert ops.packbits(torch.zeros(0, dtype=torch.uint8)).numel() == 0
    with pytest.raises(pkg.PlaidError):
        ops.packbits(torch.ones(12, dtype=torch.uint8))
    r = 0.1 * torch.randn(50, 128, generator=g)
    assert torch.equal(codec.binarize(r).cpu(), po.codec_binarize(r, sx.bucket_cutoffs, nbits))


def _ref_gpu_ops():
    from oracle import build_ref
    import os
    if not all(os.path.exists(build_ref.ref_so_path(n)) for n in build_ref.GPU_SOURCES):
        pytest.skip("oracle/_ref GPU operators not built (python oracle/build_ref.py --gpu, needs /root/reference)")
    return {n: build_ref.load(n) for n in build_ref.GPU_SOURCES}


@pytest.mark.parametrize("nbits", [1, 2, 4, 8])
def test_gpu_form_codec_operators_equal_the_reference_cuda_kernels(P, nbits):
    """The reference's OWN CUDA operators (CB/indexing/codecs/decompress_residuals.cu:8-75, packbits.cu:10-57, compiled
    from /root/reference into oracle/_ref by oracle/build_ref.py --gpu) run on this GPU: our ResidualCodec operators
    and the oracle's restatement of the GPU-form decompression must equal them bit for bit."""
    pkg, ops = P
    ref = _ref_gpu_ops()
    from reranking_multimodal_retrievers_b200.synthetic import make_synthetic_index
    sx = make_synthetic_index(300, 1, 70, nbits, seed=90 + nbits, num_centroids=512, mode="codes")
    rbm, lut = po.codec_tables(nbits)
    dev = torch.device("cuda", 0)                       # the reference kernel allocates its output on cuda:0
    bw_h, cent_h = sx.bucket_weights.half().to(dev), sx.centroids.half().to(dev)
    rbm_d, lut_d = rbm.to(dev), lut.to(dev)
    for m in (sx.num_embeddings, 1, 33, 1 << 15):       # 1 << 15: the batch size ResidualCodec.decompress splits into (residual.py:246)
        m = min(m, sx.num_embeddings)
        res, codes = sx.residuals[:m].contiguous().to(dev), sx.codes[:m].contiguous().to(dev)
        want = ref["decompress_residuals_gpu_cpp"].decompress_residuals_cpp(res, bw_h, rbm_d, lut_d, codes, cent_h, 128, nbits)
        torch.cuda.synchronize()
        assert want.dtype == torch.float16 and want.shape == (m, 128)
        ours = pkg.ResidualCodec.decompress_residuals(res, bw_h, rbm_d, lut_d, codes, cent_h, 128, nbits)
        assert torch.equal(ours.view(torch.int16), want.view(torch.int16)), "GPU-form decompression differs from the reference kernel"
        restated = po.codec_decompress_gpu_form(sx.bucket_weights, rbm, lut, sx.residuals[:m], sx.codes[:m], sx.centroids)
        assert torch.equal(restated.view(torch.int16), want.cpu().view(torch.int16)), "oracle restatement differs from the reference kernel"
        # ResidualCodec.decompress = F.normalize(...).half() of those rows on the GPU (residual.py:272-273)
        Dn_ref = torch.nn.functional.normalize(want, p=2, dim=-1).half()
        Dn_restated = po.codec_decompress_gpu_form(sx.bucket_weights, rbm, lut, sx.residuals[:m], sx.codes[:m], sx.centroids,
                                                   normalize=True)
        assert (Dn_restated.float() - Dn_ref.cpu().float()).abs().max() <= 2 ** -10
        assert (Dn_restated != Dn_ref.cpu()).float().mean() < 0.01
        Dn_ours = ops.codec_decompress_residuals(res, bw_h, rbm_d, lut_d, codes, cent_h, 128, nbits, normalize=True)
        assert (Dn_ours.float() - Dn_ref.float()).abs().max() <= 2 ** -10
    # the derived weight table is cached per tensor objects + versions (ops.build_weight_table): an in-place change of
    # the caller's bucket weights must show in the next call
    res, codes = sx.residuals[:64].contiguous().to(dev), sx.codes[:64].contiguous().to(dev)
    for _ in range(2):
        bw_h.mul_(1.5)
        want = ref["decompress_residuals_gpu_cpp"].decompress_residuals_cpp(res, bw_h, rbm_d, lut_d, codes, cent_h, 128, nbits)
        ours = pkg.ResidualCodec.decompress_residuals(res, bw_h, rbm_d, lut_d, codes, cent_h, 128, nbits)
        again = pkg.ResidualCodec.decompress_residuals(res, bw_h, rbm_d, lut_d, codes, cent_h, 128, nbits)   # cache hit
        assert torch.equal(ours.view(torch.int16), want.view(torch.int16)) and torch.equal(again, ours)
    # packbits: the reference kernel packs 32 flags per warp (sizes are multiples of 32 at its call site, residual.py:198)
    g = torch.Generator().manual_seed(100 + nbits)
    for n in (32, 128 * nbits * 7, 32 * 4097):
        flags = torch.randint(0, 2, (n,), generator=g, dtype=torch.uint8)
        flags[::5] *= 3                                   # any non-zero byte is a set flag (ballot predicate)
        want = ref["packbits_gpu_cpp"].packbits_cpp(flags.to(dev))
        torch.cuda.synchronize()
        assert torch.equal(pkg.ResidualCodec.packbits(flags.to(dev)), want)
        assert torch.equal(po.codec_packbits(flags), want.cpu())


def test_lookup_eids_and_embedding_ids_to_pids(P, golden):
    """IndexScorer.lookup_eids / embedding_ids_to_pids (CB/search/index_storage.py:61-62,82-84)."""
    pkg, ops = P
    from plaid_test_helpers import golden_host_index
    g = golden
    ix = golden_oracle_index(g)
    scorer = pkg.IndexScorer(golden_host_index(g))
    gen = torch.Generator().manual_seed(5)
    eids = torch.randint(0, ix.codes.numel(), (257,), generator=gen)
    D = scorer.lookup_eids(eids)
    rbm, lut = po.codec_tables(ix.nbits)
    want = po.codec_decompress_gpu_form(ix.bucket_weights, rbm, lut, ix.residuals[eids], ix.codes[eids],
                                        ix.centroids.half(), normalize=True)
    assert D.dtype == torch.float16 and (D.cpu().float() - want.float()).abs().max() <= 2 ** -10
    # same rows as lookup_pids (fp32 path) up to half precision
    tok2pid = torch.repeat_interleave(torch.arange(ix.doclens.numel()), ix.doclens)
    pids = scorer.embedding_ids_to_pids(eids)
    assert sorted(pids.cpu().tolist()) == sorted(set(tok2pid[eids].tolist()))
    Dp, lens = scorer.lookup_pids(torch.tensor([3, 0]))
    first = int(ix.offsets[3])
    De = scorer.lookup_eids(torch.arange(first, first + int(lens[0])))
    assert (Dp[: int(lens[0])].cpu() - De.cpu().float()).abs().max() <= 2 ** -9
    with pytest.raises(pkg.PlaidError):
        scorer.lookup_eids(torch.tensor([ix.codes.numel()]))


def test_integration_md_level2_stub(P, golden):
    """The ctypes stub INTEGRATION.md tells a maintainer to paste is executed verbatim: it must give the reference's
    filter_pids output (golden, recorded from the unmodified reference) on the reference's own S."""
    import os
    import re
    from reranking_multimodal_retrievers_b200 import build
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    block = re.search(r"level2-stub-begin.*?```python\n(.*?)```\n<!-- level2-stub-end", text, re.S).group(1)
    ns = {"PLAID_B200_LIB": build.LIB_PATH}
    exec(compile(block, "INTEGRATION.md", "exec"), ns)
    g = golden
    ix = golden_oracle_index(g)
    S, idx, cand = (torch.from_numpy(g[f"{n}_0"]) for n in ("S", "idx", "cand"))
    out = ns["filter_pids_b200"](cand, S, ix.codes, ix.doclens, ix.offsets[:-1].contiguous(), idx, int(g["ndocs"]))
    assert np.array_equal(out.cpu().numpy(), g["stage2_0"])


def test_training_scores_and_gradients(P):
    """f4: in-batch-negative scores / loss and the MaxSim backward (modeling_flmr.py:932-947,1089-1125) -- forward within
    the bf16 tolerance of the reference's recorded fp32 scores, tight against the oracle on the same bf16 operands;
    gradients against autograd through the oracle on those operands."""
    pkg, ops = P
    from plaid_test_helpers import load_golden
    g = load_golden("ib_loss")
    Q, D, mask = (torch.from_numpy(g[k]).clone() for k in ("Q", "D", "mask"))
    Qd, Dd = Q.cuda().requires_grad_(True), D.cuda().requires_grad_(True)
    s = pkg.training.in_batch_scores(Qd, Dd, mask)
    assert s.shape == (4, 12)
    ref = torch.from_numpy(g["scores"])
    assert ((s.detach().cpu() - ref).abs() <= 1e-3 * ref.abs() + 2e-2).all()            # vs the reference's fp32 scores
    Qb = Q.bfloat16().float().requires_grad_(True)
    Db = D.bfloat16().float().requires_grad_(True)
    loss_ref, s_b, _ = po.ib_loss(Qb, Db, mask)
    torch.testing.assert_close(s.detach().cpu(), s_b.detach(), rtol=2e-5, atol=2e-4)   # same operands: summation order only
    loss = pkg.training.compute_ib_loss_new(Qd, Dd, mask)
    torch.testing.assert_close(loss.detach().cpu(), loss_ref.detach(), rtol=1e-4, atol=1e-4)
    loss.backward()
    loss_ref.backward()
    # an arg-max can only differ where two passage tokens tie to within the summation order: allow a few rows
    for got, want in ((Qd.grad.cpu(), Qb.grad), (Dd.grad.cpu(), Db.grad)):
        bad = ((got - want).abs() > 1e-4 * want.abs().max() + 1e-6).any(-1).float().mean()
        assert float(bad) < 0.02, float(bad)
    assert torch.count_nonzero(Dd.grad.cpu()[~mask.squeeze(-1).bool()]) == 0           # padding never wins a maximum here
    # the paired form (FLMRModelForRetrieval.score: every query repeated for its n_docs passages)
    n_docs = int(g["n_docs"])
    Q2, D2 = Q.cuda().requires_grad_(True), D.cuda().requires_grad_(True)
    sp = pkg.training.colbert_score(Q2, D2, mask, docs_per_query=n_docs)
    torch.testing.assert_close(sp.detach().cpu(), s_b.detach().reshape(4, 4, n_docs)[torch.arange(4), torch.arange(4)].reshape(-1),
                               rtol=2e-5, atol=2e-4)
    w = torch.linspace(0.5, 1.5, sp.numel(), device="cuda")
    (sp * w).sum().backward()
    Qr = Q.bfloat16().float().requires_grad_(True)
    Dr = D.bfloat16().float().requires_grad_(True)
    sr = po.colbert_score(Qr.repeat_interleave(n_docs, dim=0), Dr, mask.squeeze(-1).bool())
    (sr * w.cpu()).sum().backward()
    for got, want in ((Q2.grad.cpu(), Qr.grad), (D2.grad.cpu(), Dr.grad)):
        bad = ((got - want).abs() > 1e-4 * want.abs().max() + 1e-6).any(-1).float().mean()
        assert float(bad) < 0.02, float(bad)
