"""Host-side containers the search path touches: a settings object, the run context, queries and
rankings.  Only the members the reference's search path reads are provided
(CB/infra/config/settings.py:11-165, CB/infra/run.py:10-60, CB/data/{queries,ranking}.py); the
training / indexing settings of the reference's ColBERTConfig are out of scope.
"""
from __future__ import annotations

import json
import os
from contextlib import contextmanager

_SEARCH_DEFAULTS = dict(
    # RunSettings (settings.py:11-60)
    root=os.path.join(os.getcwd(), "experiments"), experiment="default", nranks=1, rank=0,
    total_visible_gpus=1, index_root=None,
    # Doc/Query settings (settings.py:95-112)
    dim=128, query_maxlen=32, doc_maxlen=180, interaction="colbert",
    # IndexingSettings (settings.py:140-159)
    nbits=1, index_name=None, index_path=None,
    # SearchSettings (settings.py:161-165): None = "choose from k" in Searcher.dense_search
    ncells=None, centroid_score_threshold=None, ndocs=None,
    checkpoint=None, collection=None,
)


class ColBERTConfig:
    """Attribute bag with the reference's configure/from_existing/export semantics."""

    def __init__(self, **kw):
        self.__dict__["_assigned"] = set()
        for k, v in _SEARCH_DEFAULTS.items():
            self.__dict__[k] = v
        self.configure(**kw)

    def configure(self, **kw):
        for k, v in kw.items():
            self.__dict__[k] = v
            self._assigned.add(k)

    def __setattr__(self, k, v):
        self.configure(**{k: v})

    @property
    def index_root_(self):
        return self.index_root or os.path.join(self.root, self.experiment, "indexes/")

    @classmethod
    def from_existing(cls, *sources):
        """Later sources override earlier ones, but only with values they were explicitly given
        (CB/infra/config/base_config.py:17-31)."""
        out = cls()
        for src in sources:
            if src is None:
                continue
            if isinstance(src, dict):
                out.configure(**{k: v for k, v in src.items() if not k.startswith("_")})
            else:
                out.configure(**{k: getattr(src, k) for k in src._assigned})
        return out

    @classmethod
    def load_from_index(cls, index_path):
        """metadata.json, falling back to plan.json (base_config.py:70-87)."""
        for name in ("metadata.json", "plan.json"):
            p = os.path.join(index_path, name)
            if os.path.exists(p):
                with open(p) as f:
                    cfg = json.load(f).get("config", {})
                return cls(**{k: v for k, v in cfg.items() if isinstance(k, str)})
        raise FileNotFoundError(f"no metadata.json / plan.json under {index_path}")

    def export(self):
        return {k: v for k, v in self.__dict__.items() if not k.startswith("_")}


class RunConfig(ColBERTConfig):
    pass


class Run:
    """Singleton stack of RunConfigs (CB/infra/run.py:10-60)."""
    _instance = None

    def __new__(cls):
        if cls._instance is None:
            cls._instance = super().__new__(cls)
            cls._instance.stack = [RunConfig()]
        return cls._instance

    @property
    def config(self):
        return self.stack[-1]

    @contextmanager
    def context(self, runconfig, inherit_config=True):
        if inherit_config:
            runconfig = RunConfig.from_existing(self.config, runconfig)
        self.stack.append(runconfig)
        try:
            yield
        finally:
            self.stack.pop()


class Queries:
    """qid -> text mapping (CB/data/queries.py:13-46); the search path only uses the keys."""

    def __init__(self, path=None, data=None):
        self.path = path
        if data is None:
            raise ValueError("Queries: only in-memory `data` dicts are supported on this path")
        self.data = {qid: (c["question"] if isinstance(c, dict) else c) for qid, c in data.items()}

    @classmethod
    def cast(cls, obj):
        if isinstance(obj, cls):
            return obj
        if isinstance(obj, dict):
            return cls(data=obj)
        if isinstance(obj, (list, tuple)):
            return cls(data=dict(enumerate(obj)))
        if isinstance(obj, str):
            return cls(data={0: obj})
        raise TypeError(type(obj))

    def __len__(self):
        return len(self.data)

    def __iter__(self):
        return iter(self.data.items())

    def keys(self):
        return self.data.keys()

    def values(self):
        return self.data.values()

    def provenance(self):
        return self.path


class _LazyRows:
    """Mapping qid -> [(pid, rank, score), ...] over the arrays a batched search returns.  Rows become Python tuples
    only when they are looked at: building 1024 x 100 tuples costs several times the GPU search itself, and most
    consumers (`ranking_to_batch_results`, metrics) touch each row once or work from `Ranking.arrays()`."""

    def __init__(self, qids, pids, scores, counts):
        self._qids = list(qids)
        self._row_of = None
        self._pids, self._scores, self._counts = pids, scores, counts
        self._cache = {}

    def _row(self, i):
        row = self._cache.get(i)
        if row is None:
            n = int(self._counts[i])
            row = list(zip(self._pids[i, :n].tolist(), range(1, n + 1), self._scores[i, :n].tolist()))
            self._cache[i] = row
        return row

    def __getitem__(self, qid):
        if self._row_of is None:
            self._row_of = {q: i for i, q in enumerate(self._qids)}
        return self._row(self._row_of[qid])

    def __len__(self):
        return len(self._qids)

    def __iter__(self):
        return iter(self._qids)

    def __contains__(self, qid):
        if self._row_of is None:
            self._row_of = {q: i for i, q in enumerate(self._qids)}
        return qid in self._row_of

    def keys(self):
        return list(self._qids)

    def values(self):
        return (self._row(i) for i in range(len(self._qids)))

    def items(self):
        return ((q, self._row(i)) for i, q in enumerate(self._qids))


class Ranking:
    """qid -> [(pid, rank, score), ...] (CB/data/ranking.py:25-80).  `data` is either a plain dict (as in the
    reference) or the lazy view over result arrays that `Searcher._search_all_Q` builds (`Ranking.from_arrays`)."""

    def __init__(self, data, provenance=None):
        self.data = data
        self._provenance = provenance
        self._flat = None

    @classmethod
    def from_arrays(cls, qids, pids, scores, counts, provenance=None):
        """pids i32 / scores f32 numpy [B, k], counts [B]: rows are materialised on access."""
        return cls(_LazyRows(qids, pids, scores, counts), provenance)

    def arrays(self):
        """(qids, pids [B, k], scores [B, k], counts [B]) when built from arrays, else None."""
        d = self.data
        return (d._qids, d._pids, d._scores, d._counts) if isinstance(d, _LazyRows) else None

    @property
    def flat_ranking(self):
        if self._flat is None:
            self._flat = [(qid, *rest) for qid, sub in self.data.items() for rest in sub]
        return self._flat

    def provenance(self):
        return self._provenance

    def todict(self):
        return dict(self.data.items())

    def tolist(self):
        return list(self.flat_ranking)

    def items(self):
        return self.data.items()

    def save(self, new_path):
        assert "tsv" in os.path.basename(new_path).split("."), "rankings are saved as .tsv"
        with open(new_path, "w") as f:
            for items in self.flat_ranking:
                f.write("\t".join(str(int(x) if isinstance(x, bool) else x) for x in items) + "\n")
        with open(new_path + ".meta", "w") as f:
            json.dump({"provenance": self._provenance}, f, indent=4, default=str)
        return new_path
