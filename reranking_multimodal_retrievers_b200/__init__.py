"""B200-native (sm_100a) implementation of FLMR's ColBERT/PLAID late-interaction search path.

Public surface = the reference's: Searcher / IndexScorer, colbert_score*,
the filter_pids / decompress_residuals / segmented_maxsim / segmented_lookup operators, and the
FLMR glue create_searcher / search_custom_collection.  All arithmetic runs in libplaid_b200.so
(hand-written CUDA, C ABI in include/plaid_b200.h); there is no CPU fallback.
"""
from . import codec, training  # noqa: F401
from ._lib import PlaidError  # noqa: F401
from .codec import ResidualCodec, ResidualEmbeddings  # noqa: F401
from .infra import ColBERTConfig, Queries, Ranking, Run, RunConfig  # noqa: F401
from .modeling import (colbert_score, colbert_score_packed, colbert_score_reduce,  # noqa: F401
                       flmr_colbert_score, flmr_colbert_score_reduce)
from .ops import (codec_decompress_residuals, decompress_residuals, filter_pids, packbits,  # noqa: F401
                  segmented_lookup, segmented_maxsim)
from .search import IndexScorer, Searcher  # noqa: F401
from .searching import (create_searcher, exhaustive_search, ranking_to_batch_results,  # noqa: F401
                        search_custom_collection)
from .strided import StridedTensor  # noqa: F401

__all__ = [
    "ColBERTConfig", "Queries", "Ranking", "Run", "RunConfig", "Searcher", "IndexScorer",
    "colbert_score", "colbert_score_packed", "colbert_score_reduce", "flmr_colbert_score",
    "flmr_colbert_score_reduce", "filter_pids", "decompress_residuals", "segmented_maxsim", "segmented_lookup",
    "codec_decompress_residuals", "packbits", "ResidualCodec", "ResidualEmbeddings", "PlaidError",
    "create_searcher", "search_custom_collection", "exhaustive_search", "ranking_to_batch_results", "StridedTensor", "codec", "training",
]
