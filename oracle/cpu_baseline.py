"""CPU timing baseline ("port"): the reference's arithmetic for the bench workload, written
with the same torch CPU ops the reference dispatches to, on all host threads.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/xfmr_oracle.py header).  Used by
``bench.py``'s ``cpu_baseline`` leg and by ``bench.py --impl reference`` — the reference
itself is Python that cannot travel to the GPU box, and its verbatim (M, 1+M, D) candidate
tensor does not fit host RAM at the benchmark shape (62.9 GB per copy at M=6400),
so this is the "reference-lean" form of BASELINE.md §4: the reference's own
check_target / mask_false_negatives / loss bodies (losses.py:289-292, 483-488) applied to
``[rowdot(q,pos) | Q.Neg^T]`` logits, forward + backward to dL/dquery, fp32.

Two things to know about this port (oracle/time_reference_verbatim.py runs both on the build
container, results in profiles/cpu_reference_r01.json): (1) it is 21-43x FASTER than the
reference's own code path on the same CPU (the expand + cat of models.py:408-410 and the bmm over
the materialised tensor dominate the reference), so the reported CPU baseline is conservative;
(2) the positive logit comes from a row dot, not from the same GEMM as the pool logits, so a pool
entry that IS the row's positive does not tie bit-exactly and may survive the `<` mask - a
timing stand-in only; parity is judged against oracle/xfmr_oracle.py, which keeps exact ties.
"""

from __future__ import annotations

import time

import torch
import torch.nn.functional as F


def compute_embeds_lean(table, tokens, hist, pos, neg):
    """models.py:388-416 without the expand+cat (pos rows, shared negative pool)."""
    input_embeds = table[hist]                       # models.py:336-338
    attention_mask = (input_embeds != 0).any(-1)     # models.py:343
    query = tokens[attention_mask]                   # models.py:392
    pos_sel = pos[attention_mask]
    pos_embed = table[pos_sel]                       # models.py:400
    neg_embed = table[neg[attention_mask]]           # models.py:406
    pos_mask = pos_sel != 0                          # models.py:413
    return query[pos_mask], pos_embed[pos_mask], neg_embed


def infonce_lean(query, pos_embed, neg_embed, scale=1.0, mask_false_negatives=True):
    """losses.py:195 (as one GEMM) + :289-292 + :483-488."""
    logits = torch.cat([(query * pos_embed).sum(-1, keepdim=True), query @ neg_embed.T], dim=1)
    target = torch.zeros(logits.size(0), dtype=torch.long)
    if mask_false_negatives:
        neg_mask = logits < logits.gather(1, target[:, None])
    else:
        neg_mask = torch.ones_like(logits, dtype=torch.bool).scatter(1, target[:, None], False)
    keep = neg_mask.scatter(1, target[:, None], True)
    z = logits.where(keep, -torch.inf) * scale
    return F.cross_entropy(z, target, reduction="sum")


def train_step(table, tokens, hist, pos, neg):
    """One scoring-and-loss step: gathers + logits + InfoNCE forward + backward to dL/dtokens."""
    tokens = tokens.detach().requires_grad_(True)
    q, p, n = compute_embeds_lean(table, tokens, hist, pos, neg)
    loss = infonce_lean(q, p, n)
    loss.backward()
    return float(loss.detach()), tokens.grad


def exact_search(queries, catalog_normed, k, chunk=262144):
    """Exact cosine top-k (stable sort), the computation index.py:244-254 approximates."""
    q = F.normalize(queries, dim=-1)
    best_s = torch.full((q.size(0), 0), -torch.inf)
    best_i = torch.zeros((q.size(0), 0), dtype=torch.long)
    for lo in range(0, catalog_normed.size(0), chunk):
        s = q @ catalog_normed[lo:lo + chunk].T
        ids = torch.arange(lo, lo + s.size(1)).expand_as(s)
        s, ids = torch.cat([best_s, s], 1), torch.cat([best_i, ids], 1)
        order = torch.sort(s, dim=1, descending=True, stable=True).indices[:, :k]
        best_s, best_i = s.gather(1, order), ids.gather(1, order)
    return best_s, best_i


def time_train_steps(batch, steps, warmup=1):
    import os

    torch.set_num_threads(os.cpu_count() or 1)
    args = [torch.from_numpy(batch[k]) for k in
            ("table", "token_embeddings", "history_item_idx", "pos_item_idx", "neg_item_idx")]
    for _ in range(warmup):
        train_step(*args)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        train_step(*args)
        times.append(time.perf_counter() - t0)
    return times


def _weighted_mean(values, weights):
    """losses.py:90-111."""
    return (values * weights).sum(-1) / (weights.sum(-1) + 1e-9)


def compute_losses_lean(table, tokens, hist, pos, neg, margin=0.5, scale=1.0):
    """What ``RecommenderLightningModule.compute_losses`` evaluates per step (trainer.py:250-263):
    LogitsStatistics + all seven losses (the reference recomputes the logits for each of them: eight
    passes; here each family's logits are computed once per loss as the reference does, in the lean
    ``[rowdot | Q.Neg^T]`` form) + the backward of InfoNCE.  torch CPU ops, fp32."""
    tokens = tokens.detach().requires_grad_(True)
    q, p, n = compute_embeds_lean(table, tokens, hist, pos, neg)

    def dot_logits():
        return torch.cat([(q * p).sum(-1, keepdim=True), q @ n.T], dim=1)                 # losses.py:195

    def cos_logits():
        qn, pn, nn_ = F.normalize(q, dim=-1, eps=1e-8), F.normalize(p, dim=-1, eps=1e-8), F.normalize(n, dim=-1, eps=1e-8)
        return torch.cat([(qn * pn).sum(-1, keepdim=True), qn @ nn_.T], dim=1)            # losses.py:206-208

    def mask(l):
        return (l < l[:, :1]).float()                                                    # losses.py:289-292

    out = {}
    l = dot_logits()                                                                     # LogitsStatistics
    m = mask(l)
    negs = l[m.bool()]
    out["stats"] = (float(m.sum(1).mean()), float(l[:, 0].mean()), float(l[:, 0].std()), float(negs.mean()),
                    float(negs.std()), float(negs.min()), float(negs.max()))
    l = cos_logits(); out["AlignmentLoss"] = (1 - l[:, 0]).sum()                          # losses.py:352-353
    l = cos_logits(); m = mask(l)
    out["AlignmentContrastiveLoss"] = (1 - l[:, 0]).sum() + _weighted_mean(F.relu(l - 1 + margin), m).sum()
    l = cos_logits(); m = mask(l); out["ContrastiveLoss"] = _weighted_mean(F.relu(l - 1 + margin), m).sum()
    out["InfoNCELoss"] = infonce_lean(q, p, n, scale)                                     # losses.py:479-488
    l = dot_logits(); m = mask(l)
    out["NCELoss"] = (F.softplus(-l[:, 0]) + _weighted_mean(F.softplus(l), m)).sum()      # losses.py:498-511
    l = dot_logits(); m = mask(l)
    out["PairwiseHingeLoss"] = _weighted_mean(F.relu(l - (1 - margin) * l[:, :1]), m).sum()
    l = dot_logits(); m = mask(l)
    out["PairwiseLogisticLoss"] = _weighted_mean(F.softplus(l - (1 - margin) * l[:, :1]), m).sum()
    out["InfoNCELoss"].backward()
    return {k: (float(v.detach()) if torch.is_tensor(v) else v) for k, v in out.items()}


def time_compute_losses(batch, steps=1, warmup=0):
    import os

    torch.set_num_threads(os.cpu_count() or 1)
    args = [torch.from_numpy(batch[k]) for k in
            ("table", "token_embeddings", "history_item_idx", "pos_item_idx", "neg_item_idx")]
    for _ in range(warmup):
        compute_losses_lean(*args)
    times, last = [], None
    for _ in range(steps):
        t0 = time.perf_counter()
        last = compute_losses_lean(*args)
        times.append(time.perf_counter() - t0)
    return times, last


def time_exact_search(n_rows, n_queries, k, dim=384, seed=0):
    """Exact cosine top-k of `n_queries` queries over an `n_rows` fp32 catalog on the host cores."""
    import os

    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(seed)
    cat = F.normalize(torch.randn((n_rows, dim), generator=g), dim=-1)
    q = torch.randn((n_queries, dim), generator=g)
    t0 = time.perf_counter()
    s, i = exact_search(q, cat, k)
    return time.perf_counter() - t0
