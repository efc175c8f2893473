"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the authoring container.

The reference ships no golden vectors for its search path, so the pins are
recorded from its own executable code: a small synthetic index is written in the reference's
on-disk format, the reference ``Searcher`` (CPU branch: filter_pids.cpp, decompress_residuals.cpp,
segmented_lookup.cpp, segmented_maxsim.cpp) searches it, and every stage is tapped.

Needs /root/reference (read-only) and therefore only runs in the authoring container:

    python tests/golden/make_golden.py

Import shims (harness-side only, nothing in /root/reference is edited): a ``ujson`` alias of
``json``; ``DefaultVal.__hash__`` so the reference's config dataclasses import on Python >= 3.11;
``ColBERT.try_load_torch_extensions(False)`` which this fork never calls on the search path.
"""
import importlib
import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/third_party/ColBERT"
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    uj = types.ModuleType("ujson")
    uj.load, uj.loads = json.load, json.loads
    uj.dumps = lambda o, indent=None, **kw: json.dumps(
        o, indent=indent, default=lambda x: x.toDict() if hasattr(x, "toDict") else str(x))
    uj.dump = lambda o, f, **kw: f.write(uj.dumps(o))
    sys.modules["ujson"] = uj
    sys.path.insert(0, REF)
    os.environ.setdefault("TORCH_EXTENSIONS_DIR", "/tmp/plaid_ref_torch_ext")

    def stub(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        m.__package__ = name
        sys.modules[name] = m

    def run_init(name, path):
        spec = importlib.util.spec_from_file_location(name, f"{path}/__init__.py",
                                                      submodule_search_locations=[path])
        m = sys.modules[name]
        m.__spec__, m.__file__ = spec, spec.origin
        spec.loader.exec_module(m)

    for n, p in [("colbert", "colbert"), ("colbert.infra", "colbert/infra"),
                 ("colbert.infra.config", "colbert/infra/config")]:
        stub(n, f"{REF}/{p}")
    cc = importlib.import_module("colbert.infra.config.core_config")
    cc.DefaultVal.__hash__ = lambda self: id(self)
    run_init("colbert.infra.config", f"{REF}/colbert/infra/config")
    run_init("colbert.infra", f"{REF}/colbert/infra")
    import colbert.indexing.codecs.residual  # noqa: F401  (before index_storage: circular import)
    from colbert.infra import ColBERTConfig, Run, RunConfig
    from colbert.searcher import Searcher
    from colbert.data import Queries
    from colbert.modeling.colbert import ColBERT, colbert_score, colbert_score_packed, colbert_score_reduce
    ColBERT.try_load_torch_extensions(False)
    return dict(ColBERTConfig=ColBERTConfig, Run=Run, RunConfig=RunConfig, Searcher=Searcher,
                Queries=Queries, ColBERT=ColBERT, colbert_score=colbert_score,
                colbert_score_packed=colbert_score_packed, colbert_score_reduce=colbert_score_reduce)


def make_case(ref, name, nbits, seed, num_passages=1000, lo=8, hi=32, C=1024, B=3, Lq=64,
              ndocs=128, k=10, zero_rows=0):
    from reranking_multimodal_retrievers_b200.synthetic import (
        make_synthetic_index, make_queries, write_reference_format)
    ix = make_synthetic_index(num_passages, lo, hi, nbits, seed=seed, num_centroids=C, mode="embed")
    Q = make_queries(ix, B, Lq, seed=seed + 1, zero_rows=zero_rows)
    with tempfile.TemporaryDirectory() as root:
        exp, iname = "golden", f"{name}.nbits={nbits}"
        path = os.path.join(root, exp, "indexes", iname)
        write_reference_format(ix, path)
        with ref["Run"]().context(ref["RunConfig"](nranks=1, rank=1, root=root, experiment=exp)):
            s = ref["Searcher"](index=iname, checkpoint=None, config=ref["ColBERTConfig"](total_visible_gpus=0))
        s.configure(ndocs=ndocs)
        ranking = s._search_all_Q(ref["Queries"](data={i: f"q{i}" for i in range(B)}), Q, k=k,
                                  progress=False, remove_zero_tensors=True)
        rk = ranking.todict()
        out = dict(
            nbits=np.int64(nbits), ndocs=np.int64(ndocs), k=np.int64(k),
            ncells=np.int64(s.config.ncells), threshold=np.float64(s.config.centroid_score_threshold),
            centroids=ix.centroids.numpy(), bucket_cutoffs=ix.bucket_cutoffs.numpy(),
            bucket_weights=ix.bucket_weights.numpy(), codes=ix.codes.numpy(),
            residuals=ix.residuals.numpy(), doclens=ix.doclens.numpy(), ivf=ix.ivf.numpy(),
            ivf_lengths=ix.ivf_lengths.numpy(), Q=Q.numpy(),
            reversed_bit_map=s.ranker.codec.reversed_bit_map.numpy(),
            lookup_table=s.ranker.codec.decompression_lookup_table.numpy())
        for b in range(B):
            q = Q[b]
            q = q[torch.abs(q).sum(-1) > 0].unsqueeze(0)
            with torch.inference_mode():
                pids, S = s.ranker.retrieve(s.config, q)
                idx = S.max(-1).values >= s.config.centroid_score_threshold
                scores, fpids = s.ranker.score_pids(s.config, q, pids, S)
            out[f"cand_{b}"] = pids.numpy()
            out[f"stage2_{b}"] = fpids.numpy()
            out[f"scores_{b}"] = scores.numpy()
            out[f"rank_pids_{b}"] = np.array([p for p, _, _ in rk[b]], dtype=np.int32)
            out[f"rank_scores_{b}"] = np.array([sc for _, _, sc in rk[b]], dtype=np.float32)
            if b < 2:
                out[f"S_{b}"] = S.numpy()
                out[f"idx_{b}"] = idx.numpy()
            if b == 0:
                D = type(s.ranker).decompress_residuals(
                    fpids, s.ranker.doclens, s.ranker.embeddings_strided.codes_strided.offsets,
                    s.ranker.codec.bucket_weights, s.ranker.codec.reversed_bit_map,
                    s.ranker.codec.decompression_lookup_table, s.ranker.embeddings.residuals,
                    s.ranker.embeddings.codes, s.ranker.codec.centroids, s.ranker.codec.dim,
                    s.ranker.codec.nbits)
                out["D_0"] = D.numpy()
    # padded colbert_score (the -9999 semantics, colbert.py:268-286) on a ragged batch
    g = torch.Generator().manual_seed(seed + 2)
    n, Ld = 6, 24
    Dp = torch.nn.functional.normalize(torch.randn(n, Ld, 128, generator=g), dim=-1)
    lens = torch.tensor([24, 1, 7, 16, 3, 24])
    mask = (torch.arange(Ld).unsqueeze(0) < lens.unsqueeze(1))
    Dp = Dp * mask.unsqueeze(-1)
    Qp = torch.nn.functional.normalize(torch.randn(1, 40, 128, generator=g), dim=-1)
    out["cs_Q"], out["cs_D"], out["cs_mask"] = Qp.numpy(), Dp.numpy(), mask.numpy()
    out["cs_scores"] = ref["colbert_score"](Qp, Dp.clone(), mask, config=s.config).numpy()
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(name, {k_: (v.shape if hasattr(v, "shape") else v) for k_, v in out.items() if not k_.startswith("res")})


if __name__ == "__main__":
    torch.set_num_threads(4)
    ref = import_reference()
    make_case(ref, "plaid_nbits2", nbits=2, seed=11)
    make_case(ref, "plaid_nbits4", nbits=4, seed=23, num_passages=600, B=2, zero_rows=5)
