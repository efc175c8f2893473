"""FLMR <-> search glue with the reference's function names and arguments
(src/models/flmr/searching.py:15-63; called from src/executors/FLMR_base_executor.py:895-911)."""
from __future__ import annotations

from typing import Dict, Iterable, List, Mapping, Optional, Sequence

import torch

from . import modeling, ops
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


def exhaustive_search(query_embeddings: torch.Tensor, item_embeddings: torch.Tensor, item_embedding_mask: torch.Tensor,
                      max_K: int, truncate_scores: bool = True) -> Dict[int, list]:
    """The executor's index-free branch (FLMR_base_executor.py:918-990): every query against every item with the
    padded MaxSim (`model.score` = colbert_score, -9999 on padding, no clamp), then the best max_K per query.
    Returns {query_index: [(item_index, rank, score), ...]} like the reference, which stores `int(score)` there
    (truncate_scores=True keeps that quirk).  One pass over the items per query on the tcgen05 kernel; ties are
    ordered (score desc, item index desc) -- torch.sort leaves them unspecified."""
    Q = ops._cu(query_embeddings, torch.float32)
    D = ops._cu(item_embeddings)
    mask = ops._cu(item_embedding_mask).reshape(D.shape[0], D.shape[1])
    n_items = D.shape[0]
    K = min(int(max_K), n_items)
    out: Dict[int, list] = {}
    ids = torch.arange(n_items, device=D.device, dtype=torch.int32)
    for qi in range(Q.shape[0]):
        scores = modeling.colbert_score(Q[qi:qi + 1], D, mask, docs_per_query=n_items)
        top_ids, top_scores = ops.select_top(ids, scores, K)
        out[qi] = [(int(i), r, int(s) if truncate_scores else float(s))
                   for r, (i, s) in enumerate(zip(top_ids.tolist(), top_scores.tolist()))]
    return out


def ranking_to_batch_results(ranking_dict: Mapping, question_ids: Sequence, passage_index2id: Mapping,
                             passage_contents: Mapping, max_K: int, pos_item_ids: Optional[Iterable] = None,
                             neg_item_ids: Optional[Iterable] = None) -> List[dict]:
    """The retriever -> reranker handoff (FLMR_base_executor.py:992-1041): one record per question with
    `top_ranking_passages` = [{passage_index, passage_id, content, score}], padded to max_K by repeating the last
    entry when the index returned fewer passages (as the reference does); this is the structure
    RerankerBaseExecutor.init_retrieve reads back (Reranker_base_executor.py:244-271)."""
    rankings = list(ranking_dict.values())
    pos = list(pos_item_ids) if pos_item_ids is not None else [None] * len(rankings)
    neg = list(neg_item_ids) if neg_item_ids is not None else [None] * len(rankings)
    results = []
    for question_id, ranking_list, pos_ids, neg_ids in zip(question_ids, rankings, pos, neg):
        indices = [int(e[0]) for e in ranking_list]
        scores = [e[2] for e in ranking_list]
        if not indices:
            raise ValueError(f"question {question_id}: empty ranking (the reference would fail on it too)")
        if len(indices) < max_K:
            indices += [indices[-1]] * (max_K - len(ranking_list))
            scores += [scores[-1]] * (max_K - len(ranking_list))
        results.append({
            "question_id": question_id,
            "top_ranking_passages": [{"passage_index": i, "passage_id": passage_index2id[i],
                                      "content": passage_contents[i], "score": float(scores[n])}
                                     for n, i in enumerate(indices)],
            "pos_item_ids": pos_ids,
            "neg_item_ids": neg_ids,
        })
    return results
