"""Golden vectors for the index-build codec, recorded from the UNMODIFIED reference:
`ResidualCodec.compress` (CPU branch: fp32 `centroids @ batch.T` argmax, fp32 residual, torch.bucketize,
np.packbits; CB/indexing/codecs/residual.py:169-222) on small synthetic embeddings.

    python tests/golden/make_codec_golden.py        (authoring container only: needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, import_reference  # noqa: E402


def make_case(name, nbits, seed, n=600, C=256):
    from colbert.indexing.codecs.residual import ResidualCodec
    from colbert.infra import ColBERTConfig
    g = torch.Generator().manual_seed(seed)
    centroids = torch.nn.functional.normalize(torch.randn(C, 128, generator=g), dim=-1).half()
    assign = torch.randint(0, C, (n,), generator=g)
    embs = torch.nn.functional.normalize(centroids[assign].float() + 0.05 * torch.randn(n, 128, generator=g), dim=-1)
    res = embs - centroids[assign].float()
    qs = torch.arange(1, 2 ** nbits) / (2 ** nbits)
    cutoffs = res.flatten().quantile(qs)
    wq = (torch.arange(0, 2 ** nbits) + 0.5) / (2 ** nbits)
    weights = res.flatten().quantile(wq)
    cfg = ColBERTConfig(nbits=nbits, dim=128, total_visible_gpus=0)
    codec = ResidualCodec(config=cfg, centroids=centroids, avg_residual=res.abs().mean(), bucket_cutoffs=cutoffs,
                          bucket_weights=weights)
    assert not codec.use_gpu
    out = codec.compress(embs)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), nbits=np.int64(nbits), embs=embs.numpy(),
                        centroids_f16=centroids.numpy(), bucket_cutoffs=cutoffs.numpy(), bucket_weights=weights.numpy(),
                        codes=out.codes.numpy().astype(np.int32), residuals=out.residuals.numpy())
    print(name, out.codes.shape, out.residuals.shape, out.residuals.dtype,
          "codes == planted:", float((out.codes == assign).float().mean()))


if __name__ == "__main__":
    torch.set_num_threads(4)
    import_reference()
    make_case("codec_nbits2", 2, 31)
    make_case("codec_nbits4", 4, 32, n=300, C=128)
