"""Exact full-catalog retrieval behind the ``LanceIndex`` surface of ``xfmr_rec/index.py``.

The reference's ``search`` is an approximate IVF_HNSW_PQ query in LanceDB (index.py:194-200,
244-251) or Faiss (index.py:422, 465-467).  This index computes the exact result those
structures approximate: cosine (default, index.py:47) or inner-product scores of the query
against EVERY catalog row, history ids filtered out before ranking (index.py:239-247),
top-k by (score descending, catalog row ascending), ``score = 1 - distance`` = the cosine
itself (index.py:252-254).  Index construction, FTS / scalar indexes and the Lance storage
format are out of scope.
"""

from __future__ import annotations

import json
import pathlib
from typing import Any, Literal

import numpy as np
import pydantic
import torch

from . import ops
from .params import TOP_K


class ExactIndexConfig(pydantic.BaseModel):
    """Column names follow ``LanceIndexConfig`` (index.py:23-47)."""

    id_col: str = "item_id"
    embedding_col: str | None = "embedding"
    text_col: str = "item_text"
    index_metric: Literal["dot", "cosine"] = "cosine"
    dtype: Literal["bf16", "fp32"] = "bf16"
    max_score_bytes: int = 1 << 30  # materialised-score budget of the unfused path
    max_groupmax_bytes: int = 4 << 30  # workspace budget of the fused path (survivor lists: 256 KB per query): larger query sets run in blocks
    fused: bool = True  # tensor-core filter path (xr_score_topk) when the device / dtype allow it
    store_embeddings: bool = False  # keep the ORIGINAL fp32 rows on the host for get_ids / get_id (serving)


class ExactIndex:
    def __init__(self, config: ExactIndexConfig | None = None, device=None, *, row_offset: int = 0):
        self.config = config or ExactIndexConfig()
        self.device = torch.device(device) if device is not None else None
        self.ids: list[str] | None = None
        self.id2row: dict[str, int] | None = None
        self.catalog: torch.Tensor | None = None  # (N, D), rows normalised for the cosine metric
        self.columns: dict[str, list] = {}
        self.embeddings: torch.Tensor | None = None  # original fp32 rows (host), config.store_embeddings
        self.row_offset = row_offset  # global row of local row 0 (catalog shards)

    # -- construction -------------------------------------------------------------------------
    def index_data(self, dataset, *, overwrite: bool = False) -> "ExactIndex":
        """Accepts a HuggingFace ``datasets.Dataset`` (as index.py:135-137) or any mapping of
        column name -> sequence."""
        if self.catalog is not None and not overwrite:
            return self
        cfg = self.config
        get = (lambda c: dataset[c][:]) if hasattr(dataset, "column_names") else (lambda c: dataset[c])
        names = list(dataset.column_names) if hasattr(dataset, "column_names") else list(dataset.keys())
        self.ids = [str(x) for x in get(cfg.id_col)]
        self.id2row = {k: i for i, k in enumerate(self.ids)}
        self.columns = {c: list(get(c)) for c in names if c not in (cfg.id_col, cfg.embedding_col)}
        emb = get(cfg.embedding_col)
        emb = emb if isinstance(emb, torch.Tensor) else torch.as_tensor(np.asarray(emb, dtype=np.float32))
        return self.set_catalog(emb)

    def set_catalog(self, embeddings: torch.Tensor) -> "ExactIndex":
        """Install a (N, D) embedding matrix directly (synthetic catalogs, shards)."""
        if self.device is None:
            self.device = embeddings.device if embeddings.is_cuda else torch.device(
                "cuda", torch.cuda.current_device())
        if self.config.store_embeddings:
            self.embeddings = embeddings.detach().float().cpu()
        emb = embeddings.to(self.device)
        dt = torch.bfloat16 if self.config.dtype == "bf16" else torch.float32
        if self.config.index_metric == "cosine":
            self.catalog, _ = ops.normalize_rows(emb, 1e-12, dt)
        else:
            self.catalog = emb.to(dt).contiguous()
        if self.ids is None:
            n = emb.size(0)
            self.ids = None  # ids are the global row numbers; materialised lazily
            self.id2row = None
        return self

    def __len__(self) -> int:
        return 0 if self.catalog is None else self.catalog.size(0)

    def _row_of(self, item_id) -> int | None:
        if self.id2row is not None:
            return self.id2row.get(str(item_id))
        try:
            r = int(item_id) - self.row_offset
        except (TypeError, ValueError):
            return None
        return r if 0 <= r < len(self) else None

    def _id_of(self, row: int) -> str:
        return self.ids[row] if self.ids is not None else str(row + self.row_offset)

    # -- search -------------------------------------------------------------------------------
    def _prep_queries(self, queries: torch.Tensor) -> torch.Tensor:
        cat = self.catalog
        if self.device is None:   # catalog installed directly (a view of a resident shard)
            self.device = cat.device
        q = queries.to(self.device)
        if q.dim() == 1:
            q = q[None]
        if self.config.index_metric == "cosine":
            q, _ = ops.normalize_rows(q.float(), 1e-12, cat.dtype)
        else:
            q = q.to(cat.dtype).contiguous()
        return q

    def search_batch(self, queries: torch.Tensor, exclude_rows=None, top_k: int = TOP_K, *,
                     max_exclusions: int | None = None):
        """queries (U, D) on the device; exclude_rows: per-query lists of GLOBAL catalog rows
        (or a CSR tensor pair; then ``max_exclusions`` saves a device->host read of the longest list).
        Returns (scores (U,k) fp32, rows (U,k) int64 global; -1/-inf where fewer than k rows remain)."""
        assert self.catalog is not None, "index_data / set_catalog first"
        cat = self.catalog
        q = self._prep_queries(queries)
        u, n = q.size(0), cat.size(0)
        csr = None
        if exclude_rows is not None:
            if isinstance(exclude_rows, tuple):
                csr = exclude_rows
                if max_exclusions is None:
                    max_exclusions = int((csr[0][1:] - csr[0][:-1]).max().item()) if csr[0].numel() > 1 else 0
            else:
                if max_exclusions is None:
                    max_exclusions = max((len(l) for l in exclude_rows), default=0)
                csr = ops._csr(exclude_rows, self.device)
        max_excl = int(max_exclusions or 0) if csr is not None else 0
        fused = (self.config.fused and ops.score_topk_supported(q, cat)
                 and top_k + max_excl <= ops.SCORE_TOPK_MAX_K)
        if not fused:
            return self._finish(*self._search_materialised(q, cat, csr, top_k))
        # workspace budget: query blocks (the survivor lists take 256 KB per query)
        per_q = max(1, ops.N.lib().xr_score_topk_workspace_bytes(256, n, top_k, max_excl) // 256)
        u_blk = int(min(65535, max(256, self.config.max_groupmax_bytes // per_q // 256 * 256)))
        flags = torch.zeros(1, dtype=torch.int32, device=self.device)
        parts = []
        offs_h = csr[0].tolist() if (csr is not None and u > u_blk) else None
        for lo in range(0, u, u_blk):
            hi = min(u, lo + u_blk)
            ex = csr
            if csr is not None and offs_h is not None:
                ex = (csr[0][lo:hi + 1] - offs_h[lo], csr[1][offs_h[lo]:max(offs_h[hi], offs_h[lo] + 1)])
            s, i, _ = ops.score_topk(q[lo:hi], cat, top_k, self.row_offset, ex, max_excl, flags)
            parts.append((s, i))
        s = parts[0][0] if len(parts) == 1 else torch.cat([p[0] for p in parts])
        i = parts[0][1] if len(parts) == 1 else torch.cat([p[1] for p in parts])
        if int(flags.item()):
            # the filter could not vouch for the result (survivors lost: more rows tie with or beat the
            # sample's threshold than the sub-buckets and the overflow list hold -- massively duplicated
            # rows; or the exclusions ate the survivors): the materialised scan is exact for any input
            return self._finish(*self._search_materialised(q, cat, csr, top_k))
        return s, i          # xr_filter_finalize already reports -inf / -1 where fewer than k rows remain

    @staticmethod
    def _finish(s, i):
        # excluded rows carry -inf: they are filtered out, never returned (index.py:246)
        dead = s == float("-inf")
        return s, torch.where(dead, torch.full_like(i, -1), i)

    def _search_materialised(self, q, cat, csr, top_k):
        """Scores of catalog chunks in HBM -> exclusion mask -> streaming top-k -> merge: exact for any
        input (fp32 / D != 384 catalogs, and the fallback of the filter path)."""
        u, n = q.size(0), cat.size(0)
        chunk = max(4096, min(n, self.config.max_score_bytes // (4 * max(u, 1)) // 4 * 4))
        parts_s, parts_i = [], []
        for lo in range(0, n, chunk):
            blk = cat[lo:lo + chunk]
            sc = ops.scores(q, blk)
            if csr is not None:
                ops.mask_excluded(sc, blk.size(0), csr, col_offset=self.row_offset + lo)
            ps, pi = ops.topk(sc, top_k, n=blk.size(0), col_offset=self.row_offset + lo)
            parts_s.append(ps)
            parts_i.append(pi)
        if len(parts_s) == 1:
            return parts_s[0], parts_i[0]
        return ops.topk_merge(torch.cat(parts_s, 1), torch.cat(parts_i, 1), top_k)

    def compile_search(self, n_queries: int, top_k: int = TOP_K, max_exclusions: int = 0) -> "SearchPlan":
        """The whole search for a FIXED shape ``(n_queries, top_k)`` and at most ``max_exclusions``
        excluded rows per query as one CUDA-graph replay (the per-user validation loop of
        trainer.py:293-298 and the serving path call search once per query: host dispatch, not the
        GPU, bounds them)."""
        return SearchPlan(self, n_queries, top_k, max_exclusions)

    def search(self, embedding, exclude_item_ids: list[str] | None = None, top_k: int = TOP_K):
        """``LanceIndex.search`` (index.py:214-255): one query vector in, a
        ``datasets.Dataset`` with ``item_id`` / ``score`` (+ stored columns) out, rank order."""
        import datasets

        q = torch.as_tensor(np.asarray(embedding, dtype=np.float32))
        excl = [r for r in (self._row_of(x) for x in (exclude_item_ids or [])) if r is not None]
        excl = [r + self.row_offset for r in excl]
        s, i = self.search_batch(q[None].to(self.device), [excl], top_k)
        rows = [r - self.row_offset for r in i[0].tolist() if r >= 0]
        sc = s[0].tolist()[: len(rows)]
        out: dict[str, list] = {self.config.id_col: [self._id_of(r) for r in rows]}
        for c, vals in self.columns.items():
            out[c] = [vals[r] for r in rows]
        out["_distance"] = [1.0 - x for x in sc]
        out["score"] = sc
        return datasets.Dataset.from_dict(out)

    # -- lookups (index.py:257-292) --------------------------------------------------------------
    def get_ids(self, ids: list[str]):
        import datasets

        rows = [r for r in (self._row_of(x) for x in ids) if r is not None]
        out: dict[str, list] = {self.config.id_col: [self._id_of(r) for r in rows]}
        for c, vals in self.columns.items():
            out[c] = [vals[r] for r in rows]
        if self.config.embedding_col and self.catalog is not None:
            # the reference returns whole table rows (index.py:257-273), embedding included: the original
            # rows when they were kept (store_embeddings), else the index's own (normalised / bf16) rows
            src = self.embeddings if self.embeddings is not None else self.catalog
            sel = torch.as_tensor(rows, dtype=torch.int64, device=src.device)
            out[self.config.embedding_col] = src[sel].float().cpu().tolist() if rows else []
        return datasets.Dataset.from_dict(out)

    def get_id(self, id_val: str | None) -> dict[str, Any]:
        if id_val is None:
            return {}
        result = self.get_ids([id_val])
        return {} if len(result) == 0 else result[0]

    # -- persistence ----------------------------------------------------------------------------
    def save(self, path: str) -> None:
        p = pathlib.Path(path)
        p.mkdir(parents=True, exist_ok=True)
        torch.save({"catalog": self.catalog.cpu(), "ids": self.ids, "columns": self.columns,
                    "row_offset": self.row_offset}, p / "exact_index.pt")
        (p / "config.json").write_text(json.dumps(self.config.model_dump()))

    # -- persistence in the reference's table layout ----------------------------
    # The reference keeps items as an Arrow table {item_id, item_text, embedding: fixed_size_list
    # <float32>[D]} (the dataset handed to index_data, index.py:135-176, and what LanceDB stores).
    # These two write / read exactly that layout as Parquet, so a catalog indexed by either side can
    # be opened by the other; the search structure itself needs no persistence (there is none: the
    # index IS the embedding matrix).
    def save_table(self, path: str, *, embeddings: torch.Tensor | None = None) -> None:
        """Write ``{id_col, stored columns..., embedding_col}`` as a Parquet file.  ``embeddings``:
        the ORIGINAL fp32 rows if the caller still has them; otherwise the index's own rows (unit
        norm and/or bf16-rounded for the cosine / bf16 configuration) are written."""
        import pyarrow as pa
        import pyarrow.parquet as pq

        assert self.catalog is not None, "index_data / set_catalog first"
        emb = (embeddings if embeddings is not None else self.catalog).detach().float().cpu().numpy()
        n, d = emb.shape
        ids = self.ids if self.ids is not None else [str(r + self.row_offset) for r in range(n)]
        cols = {self.config.id_col: pa.array(ids, pa.string())}
        for c, vals in self.columns.items():
            cols[c] = pa.array(vals)
        cols[self.config.embedding_col or "embedding"] = pa.FixedSizeListArray.from_arrays(
            pa.array(np.ascontiguousarray(emb).reshape(-1), pa.float32()), d)
        meta = {b"xfmr_rec_b200.config": self.config.model_dump_json().encode(),
                b"xfmr_rec_b200.row_offset": str(self.row_offset).encode()}
        pq.write_table(pa.table(cols).replace_schema_metadata(meta), path)

    @classmethod
    def load_table(cls, path: str, config: ExactIndexConfig | None = None, device=None) -> "ExactIndex":
        """Open a Parquet / Arrow items table (written by ``save_table`` or by the reference's data
        pipeline: ``item_id, item_text, embedding``) and index it."""
        import pyarrow.parquet as pq

        table = pq.read_table(path)
        meta = table.schema.metadata or {}
        if config is None and b"xfmr_rec_b200.config" in meta:
            config = ExactIndexConfig(**json.loads(meta[b"xfmr_rec_b200.config"]))
        self = cls(config, device, row_offset=int(meta.get(b"xfmr_rec_b200.row_offset", b"0")))
        cfg = self.config
        ecol = table.column(cfg.embedding_col).combine_chunks()
        d = ecol.type.list_size if hasattr(ecol.type, "list_size") else len(ecol[0])
        flat = ecol.flatten().to_numpy(zero_copy_only=False).astype(np.float32, copy=False)
        data = {c: table.column(c).to_pylist() for c in table.column_names if c != cfg.embedding_col}
        data[cfg.embedding_col] = torch.from_numpy(flat.reshape(-1, d).copy())
        return self.index_data(data)

    @classmethod
    def load(cls, path: str, device=None) -> "ExactIndex":
        p = pathlib.Path(path)
        cfg = ExactIndexConfig(**json.loads((p / "config.json").read_text()))
        blob = torch.load(p / "exact_index.pt", weights_only=False)
        self = cls(cfg, device, row_offset=blob["row_offset"])
        self.device = self.device or torch.device("cuda", torch.cuda.current_device())
        self.catalog = blob["catalog"].to(self.device)
        self.ids = blob["ids"]
        self.id2row = None if self.ids is None else {k: i for i, k in enumerate(self.ids)}
        self.columns = blob["columns"]
        return self


class SearchPlan:
    """CUDA-graph replay of ``ExactIndex.search_batch`` for a fixed ``(U, top_k, max_exclusions)``.

    ``plan(queries, exclude)``: ``queries`` (U, D) on the device; ``exclude`` = None or a CSR pair
    ``(offsets (U+1,), rows (n,))`` of device int64 tensors with at most ``max_exclusions`` rows per
    query (checked on the DEVICE: a longer list raises the plan's flag).  Returns views of the plan's
    static output buffers ``(scores (U, k), rows (U, k))`` — valid until the next call.

    ``check=True`` (default) reads the plan's flag word after the replay (one device->host copy) and
    re-runs the search through the materialised path if a survivor list overflowed, so the result is
    exact for any input.  ``check=False`` keeps the call asynchronous; the caller then asks
    :meth:`overflowed` once after a series of searches (the flag is sticky)."""

    def __init__(self, index: ExactIndex, n_queries: int, top_k: int, max_exclusions: int = 0):
        assert index.catalog is not None, "index_data / set_catalog first"
        self.index, self.u, self.k, self.max_excl = index, int(n_queries), int(top_k), int(max_exclusions)
        dev = index.catalog.device
        cat = index.catalog
        n = cat.size(0)
        probe = torch.empty((1, cat.size(1)), dtype=cat.dtype, device=dev)
        if not (ops.score_topk_supported(probe, cat) and index.config.fused):
            raise ops.N.NativeError("compile_search needs the tensor-core scoring path "
                                    "(bf16 catalog, D = 384, sm_100)")
        if self.k + self.max_excl > ops.SCORE_TOPK_MAX_K or self.u > 65535:
            raise ValueError("top_k + max_exclusions (or the number of queries) too large for one plan")
        self.q = torch.zeros((self.u, cat.size(1)), dtype=torch.float32, device=dev)
        self.offs = torch.zeros(self.u + 1, dtype=torch.int64, device=dev)
        self.rows = torch.zeros(max(1, self.u * self.max_excl), dtype=torch.int64, device=dev)
        self.flags = torch.zeros(1, dtype=torch.int32, device=dev)
        self.ws = ops.score_topk_workspace(self.u, n, self.k, self.max_excl, dev)
        self.out = (torch.empty((self.u, self.k), dtype=torch.float32, device=dev),
                    torch.empty((self.u, self.k), dtype=torch.int64, device=dev))
        self.device = dev
        with torch.cuda.device(dev):
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):      # warm-up outside capture (one-time kernel attributes)
                self._run()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(dev)
            self.flags.zero_()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out_s, self.out_i = self._run()

    def _run(self):
        idx = self.index
        cat = idx.catalog
        if idx.config.index_metric == "cosine":
            q, _ = ops.normalize_rows(self.q, 1e-12, cat.dtype)
        else:
            q = self.q.to(cat.dtype)
        csr = (self.offs, self.rows) if self.max_excl > 0 else None
        s, i, _ = ops.score_topk(q, cat, self.k, idx.row_offset, csr, self.max_excl, self.flags,
                                 out=self.out, ws=self.ws)
        return s, i

    def overflowed(self, *, clear: bool = True) -> bool:
        """True if any search since the last clear could not be served exactly by the filter path."""
        bad = bool(int(self.flags.item()))
        if bad and clear:
            self.flags.zero_()
        return bad

    def __call__(self, queries: torch.Tensor, exclude=None, *, check: bool = True):
        assert queries.shape == self.q.shape, f"plan was compiled for {tuple(self.q.shape)} queries"
        self.q.copy_(queries, non_blocking=True)
        if self.max_excl > 0:
            if exclude is None:
                self.offs.zero_()
            else:
                offs, rows = exclude
                assert offs.numel() == self.u + 1 and rows.numel() <= self.rows.numel()
                self.offs.copy_(offs, non_blocking=True)
                self.rows[: rows.numel()].copy_(rows, non_blocking=True)
        else:
            assert exclude is None, "plan was compiled with max_exclusions = 0"
        self.graph.replay()
        if check and self.overflowed():
            cat = self.index.catalog
            s, i = self.index._search_materialised(self.index._prep_queries(queries), cat, exclude, self.k)
            return ExactIndex._finish(s, i)
        return self.out_s, self.out_i
