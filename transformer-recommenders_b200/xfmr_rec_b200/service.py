"""Serving wire format of the item index: the request / response types of
``xfmr_rec/service.py:30-72`` and the ``ItemIndex`` service surface (``service.py:137-180``) on top of
:class:`~xfmr_rec_b200.index.ExactIndex`.

The reference declares ``Query`` / ``ItemQuery`` as ``bentoml.IODescriptor`` (a pydantic model that
also accepts numpy arrays on the wire) and serves them from BentoML workers.  BentoML itself is the
serving layer and out of scope; these are the same models as plain pydantic, with
the same field names, defaults and JSON shape (embeddings as float lists), so a client of the
reference's ``/search``, ``/get_id`` and ``/get_ids`` routes can talk to an ``ItemIndexService``
unchanged.  ``search_many`` is the batched entry point the reference lacks: U queries become ONE
``search_batch`` call (one tensor-core pass over the catalog) instead of U ANN round trips.
"""

from __future__ import annotations

from typing import Annotated, Any

import numpy as np
import pydantic
import torch

from .index import ExactIndex
from .params import TOP_K


def _to_array(v: Any) -> np.ndarray | None:
    if v is None:
        return None
    return np.asarray(v, dtype=np.float32)


# float32 array on the wire: accepts lists / arrays, serialises as a (nested) list of floats
NumpyArrayType = Annotated[
    np.ndarray,
    pydantic.BeforeValidator(_to_array),
    pydantic.PlainSerializer(lambda a: None if a is None else np.asarray(a).tolist(), return_type=list | None),
]


class _Wire(pydantic.BaseModel):
    model_config = pydantic.ConfigDict(arbitrary_types_allowed=True)


class Activity(pydantic.BaseModel):
    """service.py:30-32."""

    item_id: list[str]
    item_text: list[str]


class Query(_Wire):
    """service.py:35-53: embedding + search parameters."""

    embedding: NumpyArrayType | None = None
    item_ids: list[str] | None = None
    item_texts: list[str] | None = None
    input_embeds: NumpyArrayType | None = None
    exclude_item_ids: list[str] | None = None
    top_k: int = TOP_K


class UserQuery(pydantic.BaseModel):
    """service.py:56-60."""

    user_id: str = "0"
    user_text: str = ""
    history: Activity | None = None
    target: Activity | None = None


class ItemQuery(_Wire):
    """service.py:63-66."""

    item_id: str = "0"
    item_text: str = ""
    embedding: NumpyArrayType | None = None


class ItemCandidate(pydantic.BaseModel):
    """service.py:69-72: one ranked result."""

    item_id: str
    item_text: str
    score: float


class NotFound(LookupError):
    """Stands in for ``bentoml.exceptions.NotFound`` (service.py:169, 195)."""


class ItemIndexService:
    """``ItemIndex`` of service.py:137-180 over an :class:`ExactIndex` (exact search instead of the
    IVF_HNSW_PQ query; same request / response types, same error behaviour)."""

    def __init__(self, index: ExactIndex) -> None:
        self.index = index

    def _candidates(self, rows: list[int], scores: list[float]) -> list[ItemCandidate]:
        idx = self.index
        texts = idx.columns.get(idx.config.text_col)
        out = []
        for r, s in zip(rows, scores):
            local = r - idx.row_offset
            out.append(ItemCandidate(item_id=idx._id_of(local),
                                     item_text="" if texts is None else str(texts[local]), score=float(s)))
        return out

    def search(self, query: Query) -> list[ItemCandidate]:
        """service.py:151-162."""
        assert query.embedding is not None
        return self.search_many([query])[0]

    def search_many(self, queries: list[Query]) -> list[list[ItemCandidate]]:
        """All queries in one ``search_batch`` (one pass over the catalog); per-query ``top_k`` and
        exclusion lists as in :meth:`search`."""
        if not queries:
            return []
        idx = self.index
        for q in queries:
            assert q.embedding is not None
        k = max(q.top_k for q in queries)
        emb = torch.from_numpy(np.stack([np.asarray(q.embedding, dtype=np.float32) for q in queries]))
        excl = []
        for q in queries:
            rows = [r for r in (idx._row_of(x) for x in (q.exclude_item_ids or [])) if r is not None]
            excl.append([r + idx.row_offset for r in rows])
        dev = idx.device or idx.catalog.device
        s, i = idx.search_batch(emb.to(dev), excl if any(excl) else None, k)
        s, i = s.tolist(), i.tolist()
        out = []
        for q, ss, ii in zip(queries, s, i):
            keep = [(r, sc) for r, sc in zip(ii[: q.top_k], ss[: q.top_k]) if r >= 0]
            out.append(self._candidates([r for r, _ in keep], [sc for _, sc in keep]))
        return out

    def get_id(self, item_id: str) -> ItemQuery:
        """service.py:164-171."""
        result = self.index.get_id(item_id)
        if len(result) == 0:
            raise NotFound(f"item not found: {item_id = }")
        return ItemQuery.model_validate(result)

    def get_ids(self, item_ids: list[str]) -> dict[str, ItemQuery]:
        """service.py:173-180."""
        results = self.index.get_ids(item_ids).to_list()
        items = pydantic.TypeAdapter(list[ItemQuery]).validate_python(results)
        return {item.item_id: item for item in items}
