"""Shared helpers for the parity tests (importable as plain module: tests/ is put on sys.path)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def golden_oracle_index(g):
    """OracleIndex over the tensors stored in a golden file."""
    from oracle import plaid_oracle as po
    return po.OracleIndex(
        centroids=torch.from_numpy(g["centroids"]), bucket_weights=torch.from_numpy(g["bucket_weights"]),
        codes=torch.from_numpy(g["codes"]), residuals=torch.from_numpy(g["residuals"]),
        doclens=torch.from_numpy(g["doclens"]), ivf=torch.from_numpy(g["ivf"]),
        ivf_lengths=torch.from_numpy(g["ivf_lengths"]), nbits=int(g["nbits"]),
        bucket_cutoffs=torch.from_numpy(g["bucket_cutoffs"]))


def golden_host_index(g):
    """HostIndex (product-side container) over the tensors stored in a golden file."""
    from reranking_multimodal_retrievers_b200.index import HostIndex
    return HostIndex(
        centroids=torch.from_numpy(g["centroids"]).half(), bucket_cutoffs=torch.from_numpy(g["bucket_cutoffs"]),
        bucket_weights=torch.from_numpy(g["bucket_weights"]), codes=torch.from_numpy(g["codes"]).to(torch.int32),
        residuals=torch.from_numpy(g["residuals"]), doclens=torch.from_numpy(g["doclens"]).long(),
        ivf=torch.from_numpy(g["ivf"]).to(torch.int32), ivf_lengths=torch.from_numpy(g["ivf_lengths"]).long(),
        nbits=int(g["nbits"]))


def nonzero_rows(q):
    """searcher.py:124-130 for one query [Lq, dim] (torch)."""
    return q[torch.abs(q).sum(-1) > 0]
