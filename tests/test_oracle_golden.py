"""The CPU oracle (oracle/) pinned against golden vectors recorded from the UNMODIFIED reference
Searcher (tests/golden/make_golden.py ran /root/reference in the authoring container)."""
import numpy as np
import pytest
import torch

from plaid_test_helpers import golden_oracle_index, nonzero_rows
from oracle import plaid_oracle as po


def test_codec_tables_match_reference(golden):
    rbm, lut = po.codec_tables(int(golden["nbits"]))
    assert np.array_equal(rbm.numpy(), golden["reversed_bit_map"])
    assert np.array_equal(lut.numpy(), golden["lookup_table"])


def test_stagewise_parity_with_reference(golden):
    g = golden
    ix = golden_oracle_index(g)
    Q = torch.from_numpy(g["Q"])
    k = int(g["k"])
    for b in range(Q.shape[0]):
        q = nonzero_rows(Q[b])
        r = po.rank(ix, q, int(g["ncells"]), float(g["threshold"]), int(g["ndocs"]), taps=True)
        assert np.array_equal(r["candidates"].numpy(), g[f"cand_{b}"])          # integer: bit-exact
        assert np.array_equal(r["stage2_pids"].numpy(), g[f"stage2_{b}"])      # integer: bit-exact, same order
        np.testing.assert_allclose(r["scores_unsorted"].numpy(), g[f"scores_{b}"], rtol=2e-6, atol=2e-5)
        assert np.array_equal(r["pids"][:k].numpy(), g[f"rank_pids_{b}"])
        np.testing.assert_allclose(r["scores"][:k].numpy(), g[f"rank_scores_{b}"], rtol=2e-6, atol=2e-5)
        if f"S_{b}" in g:
            np.testing.assert_allclose(r["S"].numpy(), g[f"S_{b}"], rtol=0, atol=1e-6)
            assert np.array_equal(r["idx"].numpy(), g[f"idx_{b}"])
        if b == 0:
            D = po.decompress_residuals(ix, r["stage2_pids"])
            assert np.array_equal(D.numpy(), g["D_0"])                          # one fp32 add: bit-exact


def test_padded_colbert_score_matches_reference(golden):
    g = golden
    s = po.colbert_score(torch.from_numpy(g["cs_Q"]), torch.from_numpy(g["cs_D"]), torch.from_numpy(g["cs_mask"]))
    np.testing.assert_allclose(s.numpy(), g["cs_scores"], rtol=1e-6, atol=1e-5)


def test_filter_handles_fewer_candidates_than_ndocs(golden):
    """npids < ndocs: the reference C++ pops an empty heap (UB); oracle keeps min(n, keep) like the
    reference's GPU branch (index_storage.py:138-139)."""
    g = golden
    ix = golden_oracle_index(g)
    S = torch.from_numpy(g["S_0"])
    idx = torch.from_numpy(g["idx_0"])
    cand = torch.from_numpy(g["cand_0"])[:50]
    p2, (p1, s1, s2) = po.filter_pids(ix, cand, S, idx, 128, return_stages=True)
    assert p1.numel() == 50 and p2.numel() == 32
    assert set(p2.tolist()) <= set(cand.tolist())
    assert torch.all(s2[:-1] >= s2[1:])


@pytest.mark.parametrize("name", ["codec_nbits2", "codec_nbits4"])
def test_codec_oracle_matches_reference_compress(name):
    """Index-build codec (8f-3): the restatement reproduces ResidualCodec.compress bit for bit."""
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"{name}.npz"))
    embs = torch.from_numpy(g["embs"])
    cent = torch.from_numpy(g["centroids_f16"])
    codes, res = po.codec_compress(cent, torch.from_numpy(g["bucket_cutoffs"]), int(g["nbits"]), embs)
    assert torch.equal(codes, torch.from_numpy(g["codes"]))
    assert torch.equal(res, torch.from_numpy(g["residuals"]))


def test_ib_loss_oracle_equals_reference_golden():
    """Training-time in-batch scoring: the restatement against vectors recorded by executing the reference's
    compute_ib_loss_new / colbert_score_reduce (tests/golden/make_ib_golden.py), gradients included."""
    from plaid_test_helpers import load_golden
    g = load_golden("ib_loss")
    Q, D, mask = (torch.from_numpy(g[k]).clone() for k in ("Q", "D", "mask"))
    Q.requires_grad_(True)
    D.requires_grad_(True)
    loss, scores, labels = po.ib_loss(Q, D, mask)
    assert labels.tolist() == g["labels"].tolist()
    np.testing.assert_allclose(scores.detach().numpy(), g["scores"], rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=1e-6)
    loss.backward()
    np.testing.assert_allclose(Q.grad.numpy(), g["dQ"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(D.grad.numpy(), g["dD"], rtol=1e-5, atol=1e-7)
