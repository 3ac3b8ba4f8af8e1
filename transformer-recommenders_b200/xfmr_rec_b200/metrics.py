"""Retrieval metrics — mirror of ``xfmr_rec/metrics.py`` on the device.

``compute_retrieval_metrics`` keeps the reference signature (metrics.py:17-19) and key
names (the torchmetrics function ``__name__``s, metrics.py:6-14, 76-79);
``retrieval_metrics_batch`` evaluates a whole (U, k) block of ranked lists in one kernel.
"""

from __future__ import annotations

import torch

from . import ops

METRIC_NAMES = [
    "retrieval_normalized_dcg",
    "retrieval_average_precision",
    "retrieval_auroc",
    "retrieval_precision",
    "retrieval_recall",
    "retrieval_hit_rate",
    "retrieval_reciprocal_rank",
]


def retrieval_metrics_batch(rec_idx: torch.Tensor, target_lists, top_k: int):
    """rec_idx (U,k) int64 ranked catalog rows (-1 = padding); target_lists: per-user iterables
    of catalog rows (or a CSR (offsets, ids) tensor pair).  Returns ((U,7) fp32, valid (U,) bool)
    with columns in METRIC_NAMES order; ``valid`` is False where the reference returns ``{}``."""
    return ops.retrieval_metrics(rec_idx, target_lists, top_k)


def compute_retrieval_metrics(rec_ids, target_ids, top_k: int, *, device=None):
    """xfmr_rec/metrics.py:17-79 for one ranked list of ids (any hashable ids)."""
    if len(target_ids) == 0:  # metrics.py:62-63
        return {}
    device = device or torch.device("cuda", torch.cuda.current_device())
    # ids -> dense ints; only membership matters (metrics.py:74)
    vocab: dict = {}
    targets = []
    for t in set(target_ids):
        targets.append(vocab.setdefault(t, len(vocab)))
    rec = []
    for r in list(rec_ids)[:max(top_k, len(rec_ids))]:
        rec.append(-1 if r == "" and "" not in vocab else vocab.setdefault(r, len(vocab)))
    if len(rec_ids) < top_k:  # the reference pads the caller's list in place (metrics.py:65-68)
        rec_ids += [""] * (top_k - len(rec_ids))
    rec = rec + [-1] * max(0, top_k - len(rec))
    rec_t = torch.tensor([rec], dtype=torch.int64, device=device)
    out, _ = ops.retrieval_metrics(rec_t, [targets], top_k)
    vals = out[0]
    return {name: vals[i] for i, name in enumerate(METRIC_NAMES)}
