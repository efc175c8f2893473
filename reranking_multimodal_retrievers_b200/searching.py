"""FLMR <-> search glue with the reference's function names and arguments
(src/models/flmr/searching.py:15-63; called from src/executors/FLMR_base_executor.py:895-911)."""
from __future__ import annotations

from typing import Dict

import torch

from .infra import ColBERTConfig, Queries, Run, RunConfig
from .search import Searcher


def create_searcher(index_root_path: str = ".", index_experiment_name: str = "default_experiment",
                    index_name: str = "index", use_gpu: bool = True, nbits: int = 8, device=None,
                    pid_range=None) -> Searcher:
    """Same directory convention as the reference: <root>/<experiment>/indexes/<name>.nbits=<nbits>."""
    with Run().context(RunConfig(nranks=1, rank=1, root=index_root_path, experiment=index_experiment_name)):
        total_visible_gpus = torch.cuda.device_count() if use_gpu else 0
        config = ColBERTConfig(total_visible_gpus=total_visible_gpus)
        return Searcher(index=f"{index_name}.nbits={nbits}", checkpoint=None, config=config, device=device,
                        pid_range=pid_range)


def search_custom_collection(searcher: Searcher, queries: Dict[int, str], query_embeddings: torch.Tensor,
                             num_document_to_retrieve: int = 100, remove_zero_tensors: bool = True,
                             centroid_search_batch_size: int = None):
    """Ranking {qid: [(pid, rank, score), ...]} for precomputed FLMR query embeddings [n, Lq, 128]."""
    queries = Queries(data=queries)
    return searcher._search_all_Q(queries, query_embeddings, progress=False, batch_size=centroid_search_batch_size,
                                  k=num_document_to_retrieve, remove_zero_tensors=remove_zero_tensors)
