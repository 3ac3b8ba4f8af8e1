"""CPU oracle for the scoring-and-loss + full-catalog top-k hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import it, and only as the checker / reported baseline.  The product path
(``xfmr_rec_b200``) never imports this module and fails loudly without its CUDA
library.

This is a numpy restatement (float64 arithmetic unless stated) of the reference
algorithm.  Every function cites the reference file:line it follows (paths are
relative to the upstream repository root).

Pinning status
--------------
* losses / logits statistics: PINNED — checked against the reference's own
  ``xfmr_rec/losses.py`` executed in the build container; the vectors it
  produced are committed under ``tests/golden/losses_*.npz`` together with the
  generating script ``tests/golden/make_golden.py``.
* candidate construction (``compute_embeds``): PINNED — ``xfmr_rec/models.py`` is not importable here
  (sentence_transformers absent), so ``tests/golden/make_golden_embeds.py`` executes the reference's
  own ``RecommenderModel.forward`` / ``compute_embeds`` method definitions straight from its source
  file (ast; bound to a stand-in object with a stub encoder) and stores inputs + outputs
  (``tests/golden/embeds_*.npz``); the oracle reproduces them bit for bit.
* SeqBatch construction (``seq_sample_*``): the reference draws from an unseeded generator
  (data.py:574): no value-level golden can exist.  PINNED on support and distribution:
  ``tests/golden/make_golden_seqbatch.py`` executes the reference's own ``sample_sequence`` /
  ``sample_positives`` / ``sample_negatives`` (pure numpy, taken from the source file by ast) and
  stores raw examples + marginal histograms of 6,000 draws per case; the oracle's constraints accept
  every reference example and its counter-based algorithm matches the histograms (two-sample
  chi-square).  The kernel matches the oracle's algorithm bit for bit.
* exact retrieval: the reference's search is an approximate ANN index in
  un-vendored third-party code (lancedb 0.37.1 / faiss-cpu 1.15.0); the oracle
  is the exact computation those indexes approximate — parity unpinned (and
  unpinnable) against the reference's own results.
* retrieval metrics: torchmetrics 1.9.0 is absent here; formulas restated from
  its documented behaviour — parity unpinned.
"""

from __future__ import annotations

import math

import numpy as np

LOSS_NAMES = [
    "AlignmentLoss",
    "AlignmentContrastiveLoss",
    "ContrastiveLoss",
    "InfoNCELoss",
    "NCELoss",
    "PairwiseHingeLoss",
    "PairwiseLogisticLoss",
]  # order = xfmr_rec/losses.py:546-554
COSINE_LOSSES = {"AlignmentLoss", "AlignmentContrastiveLoss", "ContrastiveLoss"}


class Config:
    """Plain stand-in for ``LossConfig`` (xfmr_rec/losses.py:11-30)."""

    def __init__(
        self,
        target_position="first",
        mask_false_negatives=True,
        num_hard_negatives=0,
        scale=1.0,
        margin=0.5,
    ):
        self.target_position = target_position
        self.mask_false_negatives = mask_false_negatives
        self.num_hard_negatives = num_hard_negatives
        self.scale = scale
        self.margin = margin


# ---------------------------------------------------------------------------
# bf16 helpers (round-to-nearest-even, as torch's .bfloat16())
# ---------------------------------------------------------------------------
def round_bf16(x: np.ndarray) -> np.ndarray:
    """Round float32 values to the nearest bf16 (ties to even), return float32."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    nan = np.isnan(x)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = rounded.view(np.float32).copy()
    out[nan] = np.nan
    return out.reshape(x.shape)


# ---------------------------------------------------------------------------
# candidate construction — xfmr_rec/models.py:388-419
# ---------------------------------------------------------------------------
def gather_rows(table: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """``nn.Embedding.forward`` (models.py:336-338, 400, 406): bit-exact row copy."""
    return table[idx]


def attention_mask_from_embeds(embeds: np.ndarray) -> np.ndarray:
    """models.py:343 — ``(input_embeds != 0).any(-1)``."""
    return (embeds != 0).any(-1)


def compute_embeds(table, token_embeddings, history_idx, pos_idx, neg_idx, *, dense=True,
                   is_normalized=False):
    """Restatement of ``RecommenderModel.compute_embeds`` (models.py:388-419).

    ``token_embeddings`` (B, L, D) stands in for the encoder output
    (models.py:345).  With ``dense=False`` the O(M^2 D) candidate tensor of
    models.py:408-416 is not materialised; the positives and the shared
    negative pool that define it are returned instead.
    """
    input_embeds = gather_rows(table, history_idx)  # models.py:336-338
    attention_mask = attention_mask_from_embeds(input_embeds)  # models.py:343, 390
    query = token_embeddings[attention_mask]  # models.py:392
    if is_normalized:  # models.py:394-395 (torch normalize, eps 1e-12)
        nrm = np.maximum(np.linalg.norm(query, axis=-1, keepdims=True), 1e-12)
        query = query / nrm
    pos_sel = pos_idx[attention_mask]  # models.py:398
    pos_embed = gather_rows(table, pos_sel)  # models.py:400
    neg_sel = neg_idx[attention_mask]  # models.py:404
    neg_embed = gather_rows(table, neg_sel)  # models.py:406
    pos_mask = pos_sel != 0  # models.py:413
    out = {
        "query_embed": query[pos_mask],  # models.py:415
        "attention_mask": attention_mask,
        "positive_mask": pos_mask,
        "pos_embed": pos_embed[pos_mask],
        "neg_embed": neg_embed,
    }
    if dense:
        m_a = pos_embed.shape[0]
        cand = np.concatenate(
            [pos_embed[:, None, :], np.broadcast_to(neg_embed[None], (m_a,) + neg_embed.shape)],
            axis=1,
        )  # models.py:408-410
        out["candidate_embed"] = cand[pos_mask]  # models.py:416
    return out


# ---------------------------------------------------------------------------
# logits — xfmr_rec/losses.py:179-209
# ---------------------------------------------------------------------------
def dot_logits(query, cand):
    """losses.py:195 — l[i,j] = q_i . c_ij ; ``cand`` is (M,C,D)."""
    return np.einsum("md,mcd->mc", np.asarray(query, np.float64), np.asarray(cand, np.float64))


def cosine_logits(query, cand, eps=1e-8):
    """losses.py:206-208 — torch cosine_similarity: each norm clamped to eps."""
    q = np.asarray(query, np.float64)
    c = np.asarray(cand, np.float64)
    qn = q / np.maximum(np.linalg.norm(q, axis=-1, keepdims=True), eps)
    cn = c / np.maximum(np.linalg.norm(c, axis=-1, keepdims=True), eps)
    return np.einsum("md,mcd->mc", qn, cn)


def lean_logits(query, pos, neg, *, cosine=False, eps=1e-8, exact_ties=True):
    """[rowdot(q,pos) | Q.Neg^T] — equal to the dense logits of models.py:408-416 +
    losses.py:195/206 for the shared-pool candidate tensor.

    In the reference the positive and every negative of a row go through ONE bmm, so a pool
    entry that is the same item as the row's positive gets a bit-identical logit and the
    strict `<` of losses.py:292 masks it.  ``exact_ties`` keeps that property here by
    reducing every (row, candidate) pair with the same elementwise-product-then-sum routine;
    ``exact_ties=False`` uses a BLAS matmul (timing baselines at large sizes only)."""
    q = np.asarray(query, np.float64)
    p = np.asarray(pos, np.float64)
    n = np.asarray(neg, np.float64)
    if cosine:
        q = q / np.maximum(np.linalg.norm(q, axis=-1, keepdims=True), eps)
        p = p / np.maximum(np.linalg.norm(p, axis=-1, keepdims=True), eps)
        n = n / np.maximum(np.linalg.norm(n, axis=-1, keepdims=True), eps)
    if not exact_ties:
        return np.concatenate([(q * p).sum(-1, keepdims=True), q @ n.T], axis=1)
    out = np.empty((q.shape[0], 1 + n.shape[0]), np.float64)
    step = max(1, (64 << 20) // max(1, 8 * max(n.shape[0], 1) * q.shape[1]))
    for lo in range(0, q.shape[0], step):
        qs = np.ascontiguousarray(q[lo:lo + step])
        out[lo:lo + step, 0] = (qs * np.ascontiguousarray(p[lo:lo + step])).sum(-1)
        out[lo:lo + step, 1:] = (qs[:, None, :] * n[None, :, :]).sum(-1)
    return out


# ---------------------------------------------------------------------------
# EmbedLoss pipeline — xfmr_rec/losses.py:211-330
# ---------------------------------------------------------------------------
def check_target(n_rows, cfg, target=None):
    """losses.py:233-261."""
    assert target is not None or cfg.target_position is not None
    assert target is None or cfg.target_position is None
    if cfg.target_position is None:
        pass
    elif cfg.target_position == "first":
        target = np.zeros(n_rows, dtype=np.int64)
    elif cfg.target_position == "diagonal":
        target = np.arange(n_rows, dtype=np.int64)
    else:
        raise ValueError(f"invalid {cfg.target_position = }")
    target = np.asarray(target)
    assert target.ndim == 1 and target.shape[0] == n_rows
    return target.astype(np.int64)


def mask_false_negatives(logits, target, cfg):
    """losses.py:283-292 — strict ``<`` against the target logit."""
    m, c = logits.shape
    if not cfg.mask_false_negatives:
        mask = np.ones((m, c), dtype=bool)
        mask[np.arange(m), target] = False
        return mask
    t = logits[np.arange(m), target][:, None]
    return logits < t


def mine_hard_negatives(logits, mask, cfg):
    """losses.py:311-330.  ``topk(sorted=False)`` leaves the choice among tied
    logits implementation-defined; the oracle (and the CUDA path) prefer the
    lower candidate index."""
    k = cfg.num_hard_negatives
    if k <= 0 or k >= logits.shape[1]:
        return mask
    masked = np.where(mask, logits, -np.inf)
    order = np.argsort(-masked, axis=1, kind="stable")[:, :k]
    sel = np.zeros_like(mask)
    np.put_along_axis(sel, order, True, axis=1)
    return mask & sel


def _weighted_mean_rows(values, weights):
    """losses.py:110-111 with dim=1."""
    w = weights.astype(np.float64)
    den = w.sum(axis=1, keepdims=True) + 1e-9
    return (values * w / den).sum(axis=1)


def _softplus(x):
    return np.logaddexp(0.0, x)


def _sigmoid(x):
    return 0.5 * (1.0 + np.tanh(0.5 * x))


def loss_from_logits(name, logits, target, mask, cfg, *, with_grad=False):
    """The seven ``loss()`` bodies (losses.py:420-543) on given logits.

    Returns the scalar loss (sum over rows) and, if asked, dL/dlogits.
    """
    l = np.asarray(logits, np.float64)
    m, c = l.shape
    rows = np.arange(m)
    t = l[rows, target]
    g = np.zeros_like(l)
    w = mask.astype(np.float64)
    cnt = w.sum(axis=1) + 1e-9
    if name == "AlignmentLoss":  # losses.py:352-353
        loss = (1.0 - t).sum()
        g[rows, target] = -1.0
    elif name in ("ContrastiveLoss", "AlignmentContrastiveLoss"):  # losses.py:370-372
        x = l - 1.0 + cfg.margin
        loss = _weighted_mean_rows(np.maximum(x, 0.0), mask).sum()
        g = (x > 0) * w / cnt[:, None]
        if name == "AlignmentContrastiveLoss":  # losses.py:445-447
            loss = loss + (1.0 - t).sum()
            g[rows, target] += -1.0
    elif name == "InfoNCELoss":  # losses.py:483-488
        keep = mask.copy()
        keep[rows, target] = True
        z = np.where(keep, l, -np.inf) * cfg.scale
        # -inf * negative scale would flip sign in torch too; reference multiplies after where
        zmax = z.max(axis=1, keepdims=True)
        e = np.exp(z - zmax)
        lse = np.log(e.sum(axis=1)) + zmax[:, 0]
        loss = (lse - z[rows, target]).sum()
        p = e / e.sum(axis=1, keepdims=True)
        g = p.copy()
        g[rows, target] -= 1.0
        g = g * cfg.scale
        g[~keep] = 0.0
    elif name == "NCELoss":  # losses.py:501-511
        sp_neg = _softplus(l)  # BCE-with-logits, label 0
        pos_loss = _softplus(-t)  # label 1
        loss = (pos_loss + _weighted_mean_rows(sp_neg, mask)).sum()
        # the weighted mean uses nce_losses, whose target column carries the
        # label-1 form; the target is never in the mask unless masking is off
        # and then scatter(False) removed it, so label-0 form is all that counts.
        g = _sigmoid(l) * w / cnt[:, None]
        g[rows, target] += -_sigmoid(-t)
    elif name in ("PairwiseHingeLoss", "PairwiseLogisticLoss"):  # losses.py:523-543
        # the reference multiplies an fp32 tensor by the Python scalar (1 - margin): an fp32 product with the
        # scalar cast to fp32 (losses.py:527, 541).  Doing the same here keeps the hinge's kink (x > 0) where
        # the reference has it; a float64 product moves it for the handful of entries with |x| ~ 1e-9.
        tm = (t.astype(np.float32) * np.float32(1.0 - cfg.margin)).astype(np.float64)
        x = l - tm[:, None]
        if name == "PairwiseHingeLoss":
            v = np.maximum(x, 0.0)
            dv = (x > 0).astype(np.float64)
        else:
            v = _softplus(x)
            dv = _sigmoid(x)
        loss = _weighted_mean_rows(v, mask).sum()
        gx = dv * w / cnt[:, None]
        g = gx.copy()
        g[rows, target] += -(1.0 - cfg.margin) * gx.sum(axis=1)
    else:
        raise KeyError(name)
    return (float(loss), g) if with_grad else float(loss)


def logits_statistics(logits, target, mask, cfg):
    """``LogitsStatistics.loss`` (losses.py:383-405); std is unbiased (torch default)."""
    l = np.asarray(logits, np.float64)
    m, c = l.shape
    num_neg = c - 1
    if cfg.num_hard_negatives > 0:
        num_neg = min(num_neg, cfg.num_hard_negatives)
    density = (mask.sum(axis=1) / (num_neg + 1e-9)).mean() if m > 0 else float("nan")
    stats = {"logits/neg/density": float(density)}
    groups = {"pos": l[np.arange(m), target], "neg": l[mask]}
    for key, v in groups.items():
        if v.size > 0:
            stats[f"logits/{key}/mean"] = float(v.mean())
            stats[f"logits/{key}/std"] = float(v.std(ddof=1)) if v.size > 1 else float("nan")
            stats[f"logits/{key}/min"] = float(v.min())
            stats[f"logits/{key}/max"] = float(v.max())
    return stats


def embed_loss(name, query, cand, cfg, target=None, *, with_grad=False, logits_dtype=None):
    """``EmbedLoss.forward`` (losses.py:150-155) on a dense (M,C,D) candidate tensor.

    ``logits_dtype='bf16'`` models Lightning's bf16-mixed autocast for the dot
    losses: inputs rounded to bf16, logits rounded to bf16 before masking
    (SURVEY §0.6); cosine logits stay fp32 there, as torch autocasts
    cosine_similarity to fp32.

    With ``with_grad`` returns (loss, dL/dquery) — the table is frozen in the
    reference (models.py:251-253) so only dL/dquery is consumed.
    """
    q = np.asarray(query, np.float64)
    c = np.asarray(cand, np.float64)
    assert q.ndim == 2 and c.ndim == 3 and q.shape[0] == c.shape[0] and q.shape[1] == c.shape[2]
    cosine = name in COSINE_LOSSES
    if cosine:
        logits = cosine_logits(q, c)
    else:
        if logits_dtype == "bf16":
            q = round_bf16(q.astype(np.float32)).astype(np.float64)
            c = round_bf16(c.astype(np.float32)).astype(np.float64)
        logits = dot_logits(q, c)
        if logits_dtype == "bf16":
            logits = round_bf16(logits.astype(np.float32)).astype(np.float64)
        elif logits_dtype == "fp32":
            logits = logits.astype(np.float32).astype(np.float64)
    tgt = check_target(logits.shape[0], cfg, target)
    mask = mask_false_negatives(logits, tgt, cfg)
    mask = mine_hard_negatives(logits, mask, cfg)
    if not with_grad:
        return loss_from_logits(name, logits, tgt, mask, cfg)
    loss, g = loss_from_logits(name, logits, tgt, mask, cfg, with_grad=True)
    dq = grad_query_from_dlogits(g, q, c, cosine=cosine)
    return loss, dq


def grad_query_from_dlogits(g, query, cand, *, cosine=False, eps=1e-8):
    """Chain rule from dL/dlogits to dL/dquery for dense candidates."""
    q = np.asarray(query, np.float64)
    c = np.asarray(cand, np.float64)
    if not cosine:
        return np.einsum("mc,mcd->md", g, c)
    qn_raw = np.linalg.norm(q, axis=-1, keepdims=True)
    qn = np.maximum(qn_raw, eps)
    cn = c / np.maximum(np.linalg.norm(c, axis=-1, keepdims=True), eps)
    qhat = q / qn
    ghat = np.einsum("mc,mcd->md", g, cn)  # dL/dqhat
    # d(q/max(|q|,eps))/dq: projection when |q| > eps, plain 1/eps scaling otherwise
    proj = ghat - (ghat * qhat).sum(-1, keepdims=True) * qhat
    return np.where(qn_raw > eps, proj / qn, ghat / eps)


def lean_loss(name, query, pos, neg, cfg, *, with_grad=False, logits_dtype=None, eps=1e-8):
    """Same objective on the shared-pool form [rowdot | Q.Neg^T] (target first)."""
    q = np.asarray(query, np.float64)
    p = np.asarray(pos, np.float64)
    n = np.asarray(neg, np.float64)
    cosine = name in COSINE_LOSSES
    if not cosine and logits_dtype == "bf16":
        q = round_bf16(q.astype(np.float32)).astype(np.float64)
        p = round_bf16(p.astype(np.float32)).astype(np.float64)
        n = round_bf16(n.astype(np.float32)).astype(np.float64)
    logits = lean_logits(q, p, n, cosine=cosine, eps=eps)
    if not cosine and logits_dtype == "bf16":
        logits = round_bf16(logits.astype(np.float32)).astype(np.float64)
    tgt = np.zeros(q.shape[0], dtype=np.int64)
    mask = mask_false_negatives(logits, tgt, cfg)
    mask = mine_hard_negatives(logits, mask, cfg)
    if not with_grad:
        return loss_from_logits(name, logits, tgt, mask, cfg)
    loss, g = loss_from_logits(name, logits, tgt, mask, cfg, with_grad=True)
    if not cosine:
        dq = g[:, :1] * p + g[:, 1:] @ n
    else:
        qn_raw = np.linalg.norm(q, axis=-1, keepdims=True)
        qn = np.maximum(qn_raw, eps)
        qhat = q / qn
        phat = p / np.maximum(np.linalg.norm(p, axis=-1, keepdims=True), eps)
        nhat = n / np.maximum(np.linalg.norm(n, axis=-1, keepdims=True), eps)
        ghat = g[:, :1] * phat + g[:, 1:] @ nhat
        proj = ghat - (ghat * qhat).sum(-1, keepdims=True) * qhat
        dq = np.where(qn_raw > eps, proj / qn, ghat / eps)
    return loss, dq, logits, mask


# ---------------------------------------------------------------------------
# exact retrieval — semantics of xfmr_rec/index.py:47, 239-254
# ---------------------------------------------------------------------------
def exact_search(queries, catalog, top_k, exclude=None, *, metric="cosine", eps=1e-12,
                 chunk=262144, dtype=np.float32):
    """Exact top-k the reference's ANN index approximates.

    cosine metric (index.py:47), history ids filtered out before ranking
    (prefilter, index.py:239-247), score = 1 - distance = cosine similarity
    (index.py:252-254).  Ranking is a total order: score descending, ties by
    lower catalog row (north_star tie rule; stable sort).
    ``exclude`` is a list (one per query) of catalog row arrays.
    """
    q = np.asarray(queries, dtype)
    if q.ndim == 1:
        q = q[None]
    cat = np.asarray(catalog, dtype)
    if metric == "cosine":
        q = q / np.maximum(np.linalg.norm(q, axis=-1, keepdims=True), eps)
    u, n = q.shape[0], cat.shape[0]
    best_s = np.full((u, 0), -np.inf, dtype)
    best_i = np.zeros((u, 0), np.int64)
    for lo in range(0, n, chunk):
        blk = cat[lo:lo + chunk]
        if metric == "cosine":
            blk = blk / np.maximum(np.linalg.norm(blk, axis=-1, keepdims=True), eps)
        s = q @ blk.T
        ids = np.broadcast_to(np.arange(lo, lo + blk.shape[0])[None], s.shape)
        if exclude is not None:
            for r in range(u):
                ex = np.asarray(exclude[r], np.int64)
                ex = ex[(ex >= lo) & (ex < lo + blk.shape[0])] - lo
                s[r, ex] = -np.inf
        s = np.concatenate([best_s, s], axis=1)
        ids = np.concatenate([best_i, ids], axis=1)
        order = np.argsort(-s, axis=1, kind="stable")[:, :top_k]  # ids ascending within ties
        best_s = np.take_along_axis(s, order, axis=1)
        best_i = np.take_along_axis(ids, order, axis=1)
    return best_s, best_i


def topk_rows(scores, k):
    """Stable top-k over a materialised (U,N) score matrix (score desc, index asc)."""
    s = np.asarray(scores)
    order = np.argsort(-s, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(s, order, axis=1), order.astype(np.int64)


# ---------------------------------------------------------------------------
# retrieval metrics — xfmr_rec/metrics.py:62-79 + torchmetrics 1.9.0 functional
# ---------------------------------------------------------------------------
METRIC_NAMES = [
    "retrieval_normalized_dcg",
    "retrieval_average_precision",
    "retrieval_auroc",
    "retrieval_precision",
    "retrieval_recall",
    "retrieval_hit_rate",
    "retrieval_reciprocal_rank",
]  # order = metrics.py:6-14


def retrieval_metrics(rec_ids, target_ids, top_k):
    """metrics.py:62-79.  Pure-Python restatement (small cases only).

    ``preds = linspace(1, 0, len(all_items))`` is strictly decreasing, so the
    ranking is the order of ``all_items`` and every metric reduces to a
    function of the hit vector over the first ``top_k`` entries and the number
    of targets.
    """
    if len(target_ids) == 0:  # metrics.py:62-63
        return {}
    recs = list(rec_ids)
    if len(recs) < top_k:  # metrics.py:65-68
        recs = recs + [""] * (top_k - len(recs))
    tset = set(target_ids)
    all_items = recs + list(tset - set(rec_ids))  # metrics.py:72
    target = [item in tset for item in all_items]  # metrics.py:74
    n_total = sum(target)
    k = min(top_k, len(all_items))
    hits = target[:k]
    n_hit = sum(hits)
    out = {}
    # nDCG@k: gains are 0/1; ideal ranking puts all n_total targets first
    dcg = sum(h / math.log2(r + 2) for r, h in enumerate(hits))
    idcg = sum(1.0 / math.log2(r + 2) for r in range(min(n_total, k)))
    out["retrieval_normalized_dcg"] = dcg / idcg if idcg > 0 else 0.0
    # AP@k: mean over hits in the top-k of (hit ordinal / position)
    if n_hit == 0:
        ap = 0.0
    else:
        acc, seen = 0.0, 0
        for r, h in enumerate(hits):
            if h:
                seen += 1
                acc += seen / (r + 1)
        ap = acc / n_hit
    out["retrieval_average_precision"] = ap
    # AUROC over the top-k: 0 when only one class is present
    n_miss = k - n_hit
    if n_hit == 0 or n_miss == 0:
        auroc = 0.0
    else:
        pairs, misses_after = 0, 0
        for h in reversed(hits):
            if h:
                pairs += misses_after
            else:
                misses_after += 1
        auroc = pairs / (n_hit * n_miss)
    out["retrieval_auroc"] = auroc
    out["retrieval_precision"] = n_hit / top_k if n_total > 0 else 0.0
    out["retrieval_recall"] = n_hit / n_total if n_total > 0 else 0.0
    out["retrieval_hit_rate"] = 1.0 if n_hit > 0 else 0.0
    rr = 0.0
    for r, h in enumerate(hits):
        if h:
            rr = 1.0 / (r + 1)
            break
    out["retrieval_reciprocal_rank"] = rr
    return out


# ---------------------------------------------------------------------------
# synthetic inputs — SURVEY §8(d); mimics data.py:669-805 (SeqBatch contract)
# ---------------------------------------------------------------------------
def synth_batch(n_items, batch, seq_len, dim=384, seed=0, pos_pad_frac=0.05, table=None):
    rng = np.random.default_rng(seed)
    if table is None:
        table = (rng.standard_normal((n_items + 1, dim)) / math.sqrt(dim)).astype(np.float32)
        table[0] = 0.0
    lens = rng.integers(1, seq_len + 1, size=batch)
    hist = np.zeros((batch, seq_len), np.int64)
    pos = np.zeros((batch, seq_len), np.int64)
    neg = np.zeros((batch, seq_len), np.int64)
    for b in range(batch):  # right-padded with 0 like pad_sequence (data.py:801)
        n = int(lens[b])
        hist[b, :n] = rng.integers(1, n_items + 1, size=n)
        pos[b, :n] = rng.integers(1, n_items + 1, size=n)
        neg[b, :n] = rng.integers(1, n_items + 1, size=n)
        drop = rng.random(n) < pos_pad_frac  # data.py:710-721: no future positive
        pos[b, :n][drop] = 0
    tokens = (rng.standard_normal((batch, seq_len, dim)) / math.sqrt(dim)).astype(np.float32)
    return {"table": table, "history_item_idx": hist, "pos_item_idx": pos,
            "neg_item_idx": neg, "token_embeddings": tokens}


# ---------------------------------------------------------------------------
# SeqBatch construction — SeqDataset.__getitem__ + collate, data.py:669-805
# (SURVEY §8f rank 2).  Two layers:
#   * seq_sample_batch: the SAME counter-based algorithm as csrc/seqbatch.cu in plain Python
#     integers (Philox4x32-10 keyed by (seed, step, row)) -> bit-exact parity of the kernel;
#   * check_seq_example: the reference's support constraints for one example (which positions /
#     positives / negatives data.py:669-747 can ever return), independent of any RNG.
# The reference draws from numpy's default_rng() (unseeded, data.py:574), so no fixed stream
# exists to match: parity is support + distribution, pinned by the constraints below.
# ---------------------------------------------------------------------------
_M32 = 0xFFFFFFFF
_M64 = 0xFFFFFFFFFFFFFFFF


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011).  counter: 4 words, key: 2 words."""
    c0, c1, c2, c3 = (int(c) & _M32 for c in counter)
    k0, k1 = (int(k) & _M32 for k in key)
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _M32, p1 & _M32, ((p0 >> 32) ^ c3 ^ k1) & _M32, p0 & _M32
        k0 = (k0 + 0x9E3779B9) & _M32
        k1 = (k1 + 0xBB67AE85) & _M32
    return c0, c1, c2, c3


def _uniform_below(words, n: int) -> int:
    return ((((words[0] << 32) | words[1]) * n) >> 64)


def seq_sample_example(history_item_idx, history_label, row, n_items, max_seq_length, pos_lookahead,
                       seed, step):
    """One example of data.py:749-785 with the device algorithm's randomness.
    Returns (history[sel], positives, negatives) as int64 arrays of length seq_len."""
    hist = np.asarray(history_item_idx, np.int64)
    lab = np.asarray(history_label, bool)
    n = len(hist)
    key64 = _splitmix64(_splitmix64((seed ^ _splitmix64(step)) & _M64) ^ (row & _M64))
    key = (key64 & _M32, key64 >> 32)
    cand = max(n - 1, 0)
    L = max_seq_length
    # sample_sequence (data.py:669-689): all positions, or the L smallest random keys, sorted
    if cand <= L:
        sel = list(range(cand))
    else:
        keys = [philox4x32_10((p, 0, 0, 0), key)[0] for p in range(cand)]
        order = sorted(range(cand), key=lambda p: (keys[p], p))
        sel = sorted(order[:L])
    seq_len = len(sel)
    prefix = np.cumsum(lab.astype(np.int64))
    # sample_positives (data.py:691-721)
    positives = np.zeros(seq_len, np.int64)
    for i, idx in enumerate(sel):
        start = idx + 1
        end = min(n, start + pos_lookahead) if pos_lookahead > 0 else n
        base = int(prefix[start - 1])
        cnt = int(prefix[end - 1]) - base if end > start else 0
        if cnt > 0:
            r = _uniform_below(philox4x32_10((i, 0, 1, 0), key), cnt)
            j = int(np.searchsorted(prefix[start:end], base + 1 + r, side="left")) + start
            positives[i] = hist[j]
    # sample_negatives (data.py:723-747)
    uniq = np.unique(hist)
    n_cand = n_items - len(uniq)
    if n_cand <= 0:
        n_cand, uniq = n_items, uniq[:0]

    def draw(i, att):
        rank = _uniform_below(philox4x32_10((i, att, 2, 0), key), n_cand)
        # rank-th item (0-based) not in the history: smallest k with uniq[k] - 1 - k > rank
        k = int(np.searchsorted(uniq - 1 - np.arange(len(uniq)), rank, side="right"))
        return rank + 1 + k

    negatives = np.zeros(seq_len, np.int64)
    if n_cand < seq_len:  # replace=True (data.py:745-747)
        for i in range(seq_len):
            negatives[i] = draw(i, 0)
    else:
        taken: set[int] = set()
        attempt = [0] * seq_len
        pending = list(range(seq_len))
        while pending:
            claims: dict[int, int] = {}
            drawn = {}
            for i in pending:
                v = draw(i, attempt[i])
                drawn[i] = v
                if v not in taken and (v not in claims or i < claims[v]):
                    claims[v] = i
            nxt = []
            for i in pending:
                v = drawn[i]
                if v not in taken and claims[v] == i:
                    negatives[i] = v
                else:
                    attempt[i] += 1
                    nxt.append(i)
            taken.update(claims)
            pending = nxt
    return hist[sel] if seq_len else hist[:0], positives, negatives


def seq_sample_batch(histories, labels, rows, n_items, max_seq_length, pos_lookahead, seed, step,
                     row_hist=None):
    """Batch + collate (data.py:787-805): right-padded with 0 to ``max_seq_length`` columns."""
    B, L = len(rows), max_seq_length
    out = {k: np.zeros((B, L), np.int64) for k in ("history_item_idx", "pos_item_idx", "neg_item_idx")}
    lens = np.zeros(B, np.int32)
    for b, row in enumerate(rows):
        h = int(row_hist[row]) if row_hist is not None else int(row)
        hs, ps, ns = seq_sample_example(histories[h], labels[h], int(row), n_items, L, pos_lookahead,
                                        seed, step)
        lens[b] = len(hs)
        out["history_item_idx"][b, :len(hs)] = hs
        out["pos_item_idx"][b, :len(hs)] = ps
        out["neg_item_idx"][b, :len(hs)] = ns
    out["seq_len"] = lens
    return out


def check_seq_example(history_item_idx, history_label, hist_out, pos_out, neg_out, n_items,
                      max_seq_length, pos_lookahead):
    """Support constraints of data.py:669-747 for ONE example (arrays already stripped of padding).
    Raises AssertionError with the violated rule."""
    hist = np.asarray(history_item_idx, np.int64)
    lab = np.asarray(history_label, bool)
    n = len(hist)
    seq_len = min(max(n - 1, 0), max_seq_length)
    assert len(hist_out) == len(pos_out) == len(neg_out) == seq_len, "sequence length (data.py:683-689)"
    # the sampled history is an order-preserving subsequence of hist[:-1]; recover one embedding
    # greedily and validate positives against EVERY embedding-consistent position set lazily:
    # positions are identifiable when items are distinct; otherwise any consistent position works
    pos_sets = []
    j = 0
    for v in hist_out:
        while j < n - 1 and hist[j] != v:
            j += 1
        assert j < n - 1, "history is not a subsequence of the user's events (data.py:687-689)"
        pos_sets.append(j)
        j += 1
    if n - 1 <= max_seq_length:
        assert pos_sets == list(range(n - 1)), "short histories are taken whole (data.py:685-686)"
    all_hist = set(hist.tolist())
    for i, idx in enumerate(pos_sets):
        start = idx + 1
        end = start + pos_lookahead if pos_lookahead > 0 else None
        cands = hist[start:end][lab[start:end]]
        if len(cands) == 0:
            # with duplicate items the greedy embedding may differ from the sampled one; accept a
            # positive only if SOME occurrence of this item has it in its window
            ok = pos_out[i] == 0
            if not ok:
                for alt in np.flatnonzero(hist[:n - 1] == hist_out[i]):
                    s2 = alt + 1
                    e2 = s2 + pos_lookahead if pos_lookahead > 0 else None
                    if pos_out[i] in set(hist[s2:e2][lab[s2:e2]].tolist()):
                        ok = True
            assert ok, "positive for a position with no future positive must be 0 (data.py:710-721)"
        else:
            ok = pos_out[i] in set(cands.tolist())
            if not ok:
                for alt in np.flatnonzero(hist[:n - 1] == hist_out[i]):
                    s2 = alt + 1
                    e2 = s2 + pos_lookahead if pos_lookahead > 0 else None
                    c2 = set(hist[s2:e2][lab[s2:e2]].tolist())
                    if (pos_out[i] in c2) or (pos_out[i] == 0 and not c2):
                        ok = True
            assert ok, "positive must be a positively-labelled event in the window (data.py:712-719)"
    neg_cands = set(range(1, n_items + 1)) - all_hist
    if not neg_cands:
        neg_cands = set(range(1, n_items + 1))
    assert set(neg_out.tolist()) <= neg_cands, "negatives must avoid the history (data.py:739-742)"
    if len(neg_cands) >= seq_len:
        assert len(set(neg_out.tolist())) == seq_len, "negatives are drawn without replacement (data.py:745-747)"


def seq_example_reference_style(rng, history_item_idx, history_label, all_idx, max_seq_length, pos_lookahead):
    """data.py:669-785 written as the reference writes it (numpy Generator, Python loop over the
    sampled positions, set difference over the whole catalog) — the CPU timing baseline of the
    device sampler and a second source for distribution comparisons."""
    hist = np.asarray(history_item_idx)
    lab = np.asarray(history_label, bool)
    indices = np.arange(len(hist) - 1)                                   # data.py:683
    if len(indices) > max_seq_length:                                    # data.py:685-689
        indices = np.sort(rng.choice(indices, size=max_seq_length, replace=False))
    positives = np.zeros_like(indices)                                   # data.py:710-721
    for i, idx in enumerate(indices):
        start = idx + 1
        end = start + pos_lookahead if pos_lookahead > 0 else None
        cands = hist[start:end][lab[start:end]]
        if len(cands) > 0:
            positives[i] = rng.choice(cands)
    neg_candidates = list(all_idx - set(hist.tolist()))                  # data.py:739-747
    if len(neg_candidates) == 0:
        neg_candidates = list(all_idx)
    negatives = rng.choice(neg_candidates, len(indices), replace=len(neg_candidates) < len(indices))
    return hist[indices], positives, negatives
