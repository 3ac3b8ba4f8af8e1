"""Batched evaluation loop.

The reference validates one user at a time (trainer.py:266-316): ``validation_step`` ->
``compute_metrics`` -> ``predict_step`` -> ``recommend`` -> ``items_index.search`` (one ANN query,
history ids excluded) -> ``compute_retrieval_metrics`` (7 torchmetrics calls) -> ``log_dict``
(Lightning averages over the epoch).  :func:`evaluate_batch` does the same for U users at once,
entirely on the device: exact full-catalog top-k with each user's history excluded
(``search_batch``), the 7 metrics of every user in one kernel (``retrieval_metrics_batch``), and the
epoch means over the users that have targets (the reference logs nothing for the others,
metrics.py:62-63).  Works with an ``ExactIndex`` or a ``dist.ShardedIndex`` (catalog sharded by rows
across ranks, one NCCL all-gather per batch).
"""

from __future__ import annotations

import torch

from . import ops
from .metrics import METRIC_NAMES, retrieval_metrics_batch


def evaluate_batch(index, query_embeds: torch.Tensor, history_rows, target_rows, top_k: int,
                   stage: str = "val"):
    """query_embeds (U, D) encoder outputs (``model.encode(history)``, trainer.py:206);
    history_rows / target_rows: per-user lists of GLOBAL catalog rows (or CSR tensor pairs):
    the history is excluded from the results (trainer.py:311-315), the targets are the items with
    a positive label (trainer.py:281-283).

    Returns ``(means, per_user, valid, rec_rows)``: ``means`` = {"<stage>/<metric>": 0-dim tensor}
    averaged over users with at least one target (what Lightning's epoch-mean of ``log_dict`` gives),
    ``per_user`` (U, 7) in METRIC_NAMES order, ``valid`` (U,) bool, ``rec_rows`` (U, top_k) int64
    ranked catalog rows (-1 padding)."""
    dev = query_embeds.device
    excl = history_rows if isinstance(history_rows, tuple) or history_rows is None else ops._csr(history_rows, dev)
    tgt = target_rows if isinstance(target_rows, tuple) else ops._csr(target_rows, dev)
    _, rec_rows = index.search_batch(query_embeds, excl, top_k)
    per_user, valid = retrieval_metrics_batch(rec_rows, tgt, top_k)
    w = valid.to(per_user.dtype)
    n = w.sum()
    means_t = (per_user * w[:, None]).sum(0) / torch.clamp(n, min=1.0)
    means = {f"{stage}/{name}": means_t[i] for i, name in enumerate(METRIC_NAMES)}
    return means, per_user, valid, rec_rows


@torch.no_grad()
def evaluate_users(encoder, table: torch.Tensor, index, history_item_idx: torch.Tensor, target_rows, top_k: int,
                   *, item_row_offset: int = -1, stage: str = "val"):
    """The whole per-user validation of the reference for U users at once, on the device: ``model.encode``
    (models.py:347-364: the user's history through the sequence encoder, pooled) -> ``items_index.search``
    with the history excluded (trainer.py:311-315) -> the seven metrics (trainer.py:266-285).

    ``history_item_idx`` (U, L) int64: rows of the item table, 0 = padding (right-padded); ``table`` the
    frozen item table with its zero row 0.  ``item_row_offset`` maps a table row to its catalog row of
    ``index`` (-1 when the catalog holds the table's rows 1.. in order).  The encoder runs in eval mode (no
    dropout) and is put back into its previous mode.  Returns what :func:`evaluate_batch` returns."""
    was_training = encoder.training
    encoder.eval()
    try:
        queries = encoder(history_item_idx, table)["sentence_embedding"]
    finally:
        encoder.train(was_training)
    hist = history_item_idx      # the WHOLE history is excluded (trainer.py:311-315), also what the encoder truncated
    mask = hist != 0
    counts = mask.sum(1)
    offs = torch.zeros(hist.size(0) + 1, dtype=torch.int64, device=hist.device)
    offs[1:] = torch.cumsum(counts, 0)
    rows = hist[mask] + item_row_offset                         # GLOBAL catalog rows (search_batch contract)
    return evaluate_batch(index, queries, (offs, rows.contiguous()), target_rows, top_k, stage)
