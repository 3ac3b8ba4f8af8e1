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
