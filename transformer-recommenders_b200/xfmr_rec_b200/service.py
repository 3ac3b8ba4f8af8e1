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
:class:`ModelService` (``Model.embed``, service.py:96-134, on the B200-native sequence encoder) and
:class:`RecommendService` (``Service.recommend_with_query`` / ``recommend_with_item_id``, service.py:202-262)
complete the request path: item ids -> stored embeddings -> sequence embedding -> exact search.
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


class ModelService:
    """``Model`` of service.py:96-134 over a :class:`~xfmr_rec_b200.encoder.SeqEncoder`: ``embed`` turns the
    ``input_embeds`` of a BATCH of queries (their items' embeddings, last ``max_seq_length`` of them) into the
    pooled sentence embedding, in one encoder forward (eval mode: no dropout)."""

    def __init__(self, encoder) -> None:
        self.encoder = encoder.eval()
        self.embed_dim: int = encoder.config.hidden_size

    def max_seq_length(self) -> int:
        """service.py:108-110."""
        return self.encoder.max_seq_length

    @torch.inference_mode()
    def embed(self, queries: list[Query]) -> list[Query]:
        """service.py:112-134: pad the per-query item embeddings to one (B, L, D) batch (a query without
        ``input_embeds`` contributes one zero row, i.e. an empty sequence), mask = any(row != 0), encode, pool."""
        if not queries:
            return queries
        dev = next(self.encoder.parameters()).device
        lmax = self.max_seq_length()
        seqs = [torch.as_tensor(np.asarray(q.input_embeds, dtype=np.float32)[-lmax:]) if q.input_embeds is not None
                else torch.zeros(1, self.embed_dim) for q in queries]
        batch = torch.nn.utils.rnn.pad_sequence(seqs, batch_first=True).to(dev)          # (B, L, D), zero padded
        b, l, d = batch.shape
        # the encoder gathers rows of a table: here the "table" is the padded batch itself
        rows = torch.arange(b * l, device=dev, dtype=torch.int64).view(b, l)
        emb = self.encoder(rows, batch.reshape(b * l, d).contiguous())["sentence_embedding"].float().cpu().numpy()
        for q, e in zip(queries, emb):
            q.embedding = e
        return queries


class RecommendService:
    """The composition of service.py:202-262 (``Service``) without the BentoML plumbing: look the query's items
    up, embed the sequence, exclude what the user already has, search."""

    def __init__(self, model: ModelService, item_index: ItemIndexService) -> None:
        self.model, self.item_index = model, item_index

    def process_query(self, query: Query) -> Query:
        """service.py:222-235: item ids -> their stored embeddings (unknown ids are dropped, the last
        ``max_seq_length`` kept)."""
        if query.item_ids is None or query.input_embeds is not None:
            return query
        items = self.item_index.get_ids(query.item_ids)
        item_ids = [i for i in query.item_ids if i in items]
        query.item_ids = item_ids[-self.model.max_seq_length():]
        embs = [items[i].embedding for i in query.item_ids]
        query.input_embeds = np.stack(embs) if embs else None
        return query

    def embed_query(self, query: Query) -> Query:
        """service.py:237-245."""
        if query.input_embeds is None or query.embedding is not None:
            return query
        return self.model.embed([query])[0]

    def recommend_with_query(self, query: Query) -> list[ItemCandidate]:
        """service.py:208-220."""
        query = self.embed_query(self.process_query(query))
        query.exclude_item_ids = [*(query.exclude_item_ids or []), *(query.item_ids or [])]
        if query.embedding is None:
            return []
        return self.item_index.search(query)

    def recommend_with_queries(self, queries: list[Query]) -> list[list[ItemCandidate]]:
        """The batched form the reference lacks: ONE encoder forward and ONE catalog pass for all queries."""
        queries = [self.process_query(q) for q in queries]
        todo = [q for q in queries if q.input_embeds is not None and q.embedding is None]
        self.model.embed(todo)
        for q in queries:
            q.exclude_item_ids = [*(q.exclude_item_ids or []), *(q.item_ids or [])]
        have = [q for q in queries if q.embedding is not None]
        found = iter(self.item_index.search_many(have))
        return [next(found) if q.embedding is not None else [] for q in queries]

    def recommend_with_user_id(self, user_index: "UserIndexService", user_id: str,
                                exclude_item_ids: list[str] | None = None, top_k: int = TOP_K) -> list[ItemCandidate]:
        """service.py:264-289: the user's history and target items form the query sequence (and are excluded)."""
        user = user_index.get_id(user_id)
        item_ids: list[str] = []
        item_texts: list[str] = []
        for act in (user.history, user.target):
            if act:
                item_ids += act.item_id
                item_texts += act.item_text
        return self.recommend_with_query(Query(item_ids=item_ids, item_texts=item_texts,
                                               exclude_item_ids=exclude_item_ids, top_k=top_k))

    def recommend_with_item_id(self, item_id: str, exclude_item_ids: list[str] | None = None,
                               top_k: int = TOP_K) -> list[ItemCandidate]:
        """service.py:247-262."""
        item = self.item_index.get_id(item_id)
        query = Query(item_ids=[item.item_id], item_texts=[item.item_text],
                      input_embeds=item.embedding[None, :] if item.embedding is not None else None,
                      exclude_item_ids=exclude_item_ids, top_k=top_k)
        return self.recommend_with_query(query)


class UserIndexService:
    """``UserIndex`` of service.py:183-199: a key lookup over the users table (``user_id``, ``user_text``,
    ``history`` / ``target`` activities).  The reference keeps the table in LanceDB; any mapping or iterable of
    rows in that column layout serves (e.g. ``datasets.Dataset.to_list()`` of its users parquet)."""

    def __init__(self, rows) -> None:
        if isinstance(rows, dict):
            rows = rows.values()
        self._rows = {str(r["user_id"]): r for r in rows}

    def get_id(self, user_id: str) -> UserQuery:
        row = self._rows.get(str(user_id))
        if row is None:
            raise NotFound(f"user not found: {user_id = }")
        return UserQuery.model_validate(row)
