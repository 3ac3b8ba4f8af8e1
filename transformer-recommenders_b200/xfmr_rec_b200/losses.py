"""Drop-in mirror of ``xfmr_rec/losses.py`` backed by the xfmr_b200 CUDA kernels.

Same names, constructor (``cls(config)``), call signature
(``loss(query_embed, candidate_embed, target=None) -> 0-dim tensor``, sum over rows), error
behaviour (``AssertionError`` on shape / target misuse, ``ValueError`` on a bad
``target_position``) and registry (``LOSS_CLASSES`` order, ``LossType``) as the reference
(xfmr_rec/losses.py:11-30, 114-155, 546-564).  The modules hold no parameters or buffers, so
Lightning checkpoints keep their keys (trainer.py:352-362).

``candidate_embed`` may be
  * a dense ``(M, C, D)`` tensor — the reference's own layout (models.py:410), or
  * a :class:`PoolCandidates` handle — positives ``(M, D)`` + ONE shared negative pool
    ``(Cn, D)``, mathematically the tensor models.py:408-416 builds, without the O(M^2 D) copy, or
  * a :class:`SampledCandidates` handle — per-row candidate indices into the item table
    (BASELINE config 3).
All arithmetic runs in libxfmr_b200.so; there is no PyTorch/CPU fallback.
"""

from __future__ import annotations

import abc
from typing import Literal

import pydantic
import torch

from . import _native as N
from . import ops


class LossConfig(pydantic.BaseModel):
    """Configuration for embedding losses (xfmr_rec/losses.py:11-30)."""

    target_position: Literal["first", "diagonal"] | None = "first"
    mask_false_negatives: bool = True
    num_hard_negatives: int = 0
    scale: float = 1.0
    margin: float = 0.5


# "reference": follow what the reference computes under the ambient autocast state
#              (dot logits in bf16 under bf16-mixed autocast or for bf16 inputs, cosine in fp32
#              unless the inputs are bf16 — SURVEY §0.6);  "bf16" / "fp32": force.
_PRECISION = "reference"
_MAX_LOGIT_BYTES = 512 << 20  # row-chunk the materialised paths beyond this


def set_precision(mode: str) -> None:
    global _PRECISION
    assert mode in ("reference", "bf16", "fp32")
    _PRECISION = mode


# ---------------------------------------------------------------------------------------------
# structured candidate handles (SURVEY §8b "Candidate producer")
# ---------------------------------------------------------------------------------------------
class _CandidateHandle:
    requires_grad = False

    def dim(self) -> int:
        return 3

    def size(self, i: int | None = None):
        return self.shape if i is None else self.shape[i]


class PoolCandidates(_CandidateHandle):
    """candidate_embed[i] = [pos[i] | neg[0..Cn)] for every row i (models.py:402-410)."""

    def __init__(self, pos: torch.Tensor, neg: torch.Tensor):
        assert pos.dim() == 2 and neg.dim() == 2 and pos.size(1) == neg.size(1)
        self.pos, self.neg = pos, neg

    @property
    def shape(self):
        return torch.Size((self.pos.size(0), 1 + self.neg.size(0), self.pos.size(1)))

    @property
    def device(self):
        return self.pos.device

    @property
    def dtype(self):
        return self.pos.dtype

    def __getitem__(self, mask):  # candidate_embed[pos_mask], models.py:416
        return PoolCandidates(self.pos[mask], self.neg)

    def dense(self) -> torch.Tensor:
        """Materialise the reference's (M, 1+Cn, D) tensor (tests / tiny shapes only)."""
        m = self.pos.size(0)
        return torch.cat([self.pos[:, None, :], self.neg[None].expand(m, -1, -1)], dim=1)


_SAMPLED_ONE_PASS = True  # False: logits / rowloss / dq as three launches (kept for A/B timing and tests)


class SampledCandidates(_CandidateHandle):
    """candidate_embed[i, j] = table[cand_idx[i, j]]; column 0 is the positive."""

    def __init__(self, table: torch.Tensor, cand_idx: torch.Tensor, table_inv_norm=None):
        assert table.dim() == 2 and cand_idx.dim() == 2 and cand_idx.dtype == torch.int64
        self.table, self.cand_idx, self.table_inv_norm = table, cand_idx, table_inv_norm

    @property
    def shape(self):
        return torch.Size((*self.cand_idx.shape, self.table.size(1)))

    @property
    def device(self):
        return self.table.device

    @property
    def dtype(self):
        return self.table.dtype

    def __getitem__(self, mask):
        return SampledCandidates(self.table, self.cand_idx[mask], self.table_inv_norm)

    def dense(self) -> torch.Tensor:
        return ops.gather_rows(self.table, self.cand_idx)


class _AttachGrad(torch.autograd.Function):
    """loss value + precomputed dL/dquery -> autograd edge (forward and backward are fused in
    the kernels; backward only scales the stored gradient by grad_output)."""

    @staticmethod
    def forward(ctx, query, loss, dq):
        ctx.save_for_backward(dq)
        ctx.q_dtype = query.dtype
        return loss.view(())

    @staticmethod
    def backward(ctx, grad_out):
        (dq,) = ctx.saved_tensors
        return (dq * grad_out).to(ctx.q_dtype), None, None


class _AttachGradBoth(torch.autograd.Function):
    """Dense (M, C, D) candidate tensor that requires grad (the reference's losses are differentiable in
    both arguments, losses.py:128-155): the precomputed dL/dquery and dL/dcandidate_embed."""

    @staticmethod
    def forward(ctx, query, cand, loss, dq, dcand):
        ctx.save_for_backward(dq, dcand)
        ctx.q_dtype, ctx.c_dtype = query.dtype, cand.dtype
        ctx.q_grad = dq is not None
        return loss.view(())

    @staticmethod
    def backward(ctx, grad_out):
        dq, dcand = ctx.saved_tensors
        gq = (dq * grad_out).to(ctx.q_dtype) if ctx.needs_input_grad[0] and dq is not None else None
        return gq, (dcand * grad_out).to(ctx.c_dtype), None, None, None


class _AttachGradScatter(torch.autograd.Function):
    """Same, but connected to the encoder output the query rows were compacted from
    (models.compute_embeds): backward is ONE kernel — grad_output scale, cast and scatter into the
    (B*L, D) layout with zeros for unselected positions (autograd of models.py:392, 415)."""

    @staticmethod
    def forward(ctx, tok2d, loss, dq, inv_pos):
        ctx.save_for_backward(dq, inv_pos)
        ctx.n_rows, ctx.tok_dtype = tok2d.size(0), tok2d.dtype
        return loss.view(())

    @staticmethod
    def backward(ctx, grad_out):
        dq, inv_pos = ctx.saved_tensors
        return ops.scatter_scaled(dq, inv_pos, grad_out, ctx.n_rows, ctx.tok_dtype), None, None, None


class _OneLoss:
    """losses[kind] container of the fused path (only the requested kind was evaluated)."""

    def __init__(self, kind, value):
        self.kind, self.value = kind, value

    def __getitem__(self, k):
        return self.value if k == self.kind else torch.zeros((), device=self.value.device)


def weighted_mean(values, sample_weights, *, dim=None, keepdim=False):
    """xfmr_rec/losses.py:90-111 (kept for API parity; the kernels fuse it)."""
    denominator = sample_weights.sum(dim=dim, keepdim=True) + 1e-9
    return (values * sample_weights / denominator).sum(dim=dim, keepdim=keepdim)


def _autocast_bf16() -> bool:
    try:
        return torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16
    except TypeError:  # older signature
        return torch.is_autocast_enabled() and torch.get_autocast_gpu_dtype() == torch.bfloat16


_FUSED_KINDS = {"InfoNCELoss", "NCELoss", "PairwiseHingeLoss", "PairwiseLogisticLoss",
                "ContrastiveLoss", "AlignmentContrastiveLoss"}


class EmbedLoss(torch.nn.Module, abc.ABC):
    """Base class: check -> logits -> target -> false-negative mask -> hard negatives -> loss
    (xfmr_rec/losses.py:114-155), executed by the CUDA library."""

    COSINE = False

    def __init__(self, config: LossConfig) -> None:
        super().__init__()
        self.config = config

    # -- reference validations --------------------------------------------------------------
    def check_embeds(self, query_embed, candidate_embed) -> None:
        """xfmr_rec/losses.py:157-177."""
        assert query_embed.dim() == 2, f"{query_embed.dim() = }, {query_embed.size() = }"
        assert candidate_embed.dim() == 3, (
            f"{candidate_embed.dim() = }, {candidate_embed.size() = }"
        )
        assert query_embed.size(0) == candidate_embed.size(0), (
            f"{query_embed.size(0) = } != {candidate_embed.size(0) = }"
        )
        assert query_embed.size(-1) == candidate_embed.size(-1), (
            f"{query_embed.size(-1) = } != {candidate_embed.size(-1) = }"
        )

    def check_target(self, n_rows: int, n_cand: int, target):
        """xfmr_rec/losses.py:211-261 -> (target_mode, target tensor | None)."""
        assert target is not None or self.config.target_position is not None, (
            "either `targets` or `config.target_position` must be provided"
        )
        assert target is None or self.config.target_position is None, (
            "only one of `targets` or `config.target_position` should be provided"
        )
        match self.config.target_position:
            case None:
                assert target is not None
                assert target.dim() == 1, f"{target.dim() = }, {target.size() = }"
                assert target.size(0) == n_rows, f"{target.size(0) = } != {n_rows = }"
                return N.TARGET_EXPLICIT, target.to(torch.int64).contiguous()
            case "first":
                return N.TARGET_FIRST, None
            case "diagonal":
                if n_rows > n_cand:
                    raise IndexError("diagonal targets need num_candidates >= batch_size")
                return N.TARGET_DIAGONAL, None
            case _:
                msg = f"invalid {self.config.target_position = }"
                raise ValueError(msg)

    # -- precision policy ---------------------------------------------------------------------
    def _compute_dtype(self, query_embed) -> tuple[torch.dtype, bool]:
        inputs_bf16 = query_embed.dtype == torch.bfloat16
        if _PRECISION == "bf16":
            bf16 = True
        elif _PRECISION == "fp32":
            bf16 = False
        elif self.COSINE:
            bf16 = inputs_bf16  # torch autocasts cosine_similarity to fp32
        else:
            bf16 = inputs_bf16 or _autocast_bf16()
        # dot logits come out of an autocast bmm in bf16 (losses.py:195); cosine stay fp32
        return (torch.bfloat16 if bf16 else torch.float32), (bf16 and not self.COSINE)

    # -- execution ----------------------------------------------------------------------------
    def _evaluate(self, query_embed, candidate_embed, target, *, want_stats=False, want_dcand=False):
        """Returns (losses f64[7], stats f64[16] | None, dq fp32 | None).  ``want_dcand`` (forward() only): a
        dense candidate tensor that requires grad also gets its gradient, left in ``self._dcand`` for forward()."""
        self.check_embeds(query_embed, candidate_embed)
        if not query_embed.is_cuda:
            raise N.NativeError("xfmr_rec_b200 losses need CUDA tensors; there is no CPU fallback")
        m, c = candidate_embed.size(0), candidate_embed.size(1)
        mode, tgt = self.check_target(m, c, target)
        need_grad = torch.is_grad_enabled() and query_embed.requires_grad
        self._dcand = None
        cand_grad = (want_dcand and torch.is_grad_enabled()
                     and bool(getattr(candidate_embed, "requires_grad", False)))
        if cand_grad and not isinstance(candidate_embed, torch.Tensor):
            raise NotImplementedError(
                "gradients w.r.t. a candidate HANDLE are not produced: the item table is frozen in the "
                "reference (models.py:251-253); pass the dense (M, C, D) tensor to differentiate it"
            )
        if m == 0:  # empty batch: every sum is 0 (and the stats block reports zero rows)
            z = torch.zeros(N.XR_NUM_LOSSES, dtype=torch.float64, device=query_embed.device)
            st = torch.zeros(N.XR_STATS_SLOTS, dtype=torch.float64, device=query_embed.device)
            dq0 = torch.zeros_like(query_embed, dtype=torch.float32) if need_grad else None
            return z, (st if want_stats else None), dq0
        cdt, logits_bf16 = self._compute_dtype(query_embed)
        cfg = ops.make_cfg(self.config, logits_bf16=logits_bf16)
        kind = N.LOSS_KIND.get(type(self).__name__, -1)
        grad_kind = kind if (need_grad or cand_grad) else -1
        q = query_embed.detach()

        if isinstance(candidate_embed, _CandidateHandle) and mode != N.TARGET_FIRST:
            candidate_embed = candidate_embed.dense()  # handles define the positive as column 0

        if isinstance(candidate_embed, PoolCandidates):
            return self._pool(q, candidate_embed, cfg, cdt, kind, grad_kind, want_stats)
        if isinstance(candidate_embed, SampledCandidates):
            return self._sampled(q, candidate_embed, cfg, cdt, grad_kind, want_stats)
        return self._dense(q, candidate_embed.detach(), cfg, cdt, mode, tgt, grad_kind, want_stats, cand_grad)

    def _pool(self, q, cand, cfg, cdt, kind, grad_kind, want_stats):
        pos, neg = cand.pos.detach(), cand.neg.detach()
        if type(self).__name__ == "AlignmentLoss":
            neg = neg[:0]  # 1 - cos(q, pos): no negative is ever read (losses.py:352-353)
        q_inv = None
        if self.COSINE:  # losses.py:206-208: both sides normalised, norms clamped at 1e-8
            q, q_inv = ops.normalize_rows(q, 1e-8, cdt)
            pos, _ = ops.normalize_rows(pos, 1e-8, cdt)
            neg = ops.normalize_rows(neg, 1e-8, cdt)[0] if neg.size(0) else neg.to(cdt)
        else:
            q, pos, neg = (t.to(cdt).contiguous() for t in (q, pos, neg))
        # the fused softmax takes scale * target as its reference maximum, which needs scale > 0; the
        # reference accepts any scale (losses.py:483-488), so those configs use the materialised path
        fused_ok = (not want_stats and cfg.num_hard_negatives == 0 and neg.size(0) > 0
                    and type(self).__name__ in _FUSED_KINDS and ops.fused_pool_supported(q, neg)
                    and (type(self).__name__ != "InfoNCELoss" or cfg.scale > 0))
        if (want_stats and grad_kind < 0 and cfg.num_hard_negatives == 0 and neg.size(0) > 0
                and ops.fused_pool_supported(q, neg) and (self.COSINE or cfg.scale > 0)):
            # every loss of this logit family + the statistics block from ONE tensor-core pass
            losses, stats = ops.fused_pool_all(q, pos, neg, cfg, self.COSINE)
            return losses, stats, None
        if fused_ok:
            loss, dq, _ = ops.fused_pool_loss(q, pos, neg, kind, cfg, q_inv=q_inv,
                                              want_grad=grad_kind >= 0)
            return _OneLoss(kind, loss.view(torch.float32)[2]), None, dq
        # materialised path, row-chunked so the logits stay bounded
        m, cn = q.size(0), neg.size(0)
        rows = max(1, min(m, _MAX_LOGIT_BYTES // (4 * (cn + 4)))) if m else 1
        losses, stats, dqs = None, None, []
        for lo in range(0, max(m, 1), rows):
            qs, ps = q[lo:lo + rows], pos[lo:lo + rows]
            logits = ops.logits_pool(qs, ps, neg)
            l, s, dl = ops.rowloss(logits, cn + 1, cfg, N.TARGET_LAST, None, grad_kind,
                                   want_stats=want_stats)
            losses = l if losses is None else losses + l
            stats = _merge_stats(stats, s)
            if dl is not None:
                dqs.append(ops.dq_pool(dl, qs, ps, neg, cosine=self.COSINE,
                                       q_inv=None if q_inv is None else q_inv[lo:lo + rows]))
        dq = (torch.cat(dqs) if len(dqs) > 1 else dqs[0]) if dqs else None
        return losses, stats, dq

    def _sampled(self, q, cand, cfg, cdt, grad_kind, want_stats):
        table = cand.table if cand.table.dtype == cdt else cand.table.to(cdt)
        q = q.to(cdt).contiguous()
        idx = cand.cand_idx.contiguous()
        q_inv = table_inv = None
        if self.COSINE:
            _, q_inv = ops.normalize_rows(q, 1e-8, want_y=False)
            table_inv = cand.table_inv_norm
            if table_inv is None:
                _, table_inv = ops.normalize_rows(table, 1e-8, want_y=False)
        if q.size(1) == 384 and 1 <= idx.size(1) <= 8192 and _SAMPLED_ONE_PASS:
            # one launch: the row's logits and their gradient never leave shared memory
            return ops.sampled_step(q, table, idx, cfg, table_inv, q_inv, grad_kind,
                                    want_stats=want_stats)
        logits = ops.logits_sampled(q, table, idx, table_inv, q_inv)
        losses, stats, dl = ops.rowloss(logits, idx.size(1), cfg, N.TARGET_FIRST, None, grad_kind,
                                        want_stats=want_stats)
        dq = ops.dq_sampled(dl, q, table, idx, table_inv, q_inv) if dl is not None else None
        return losses, stats, dq

    def _dense(self, q, cand, cfg, cdt, mode, tgt, grad_kind, want_stats, cand_grad=False):
        q, cand = q.to(cdt).contiguous(), cand.to(cdt).contiguous()
        q_inv = None
        if self.COSINE:
            _, q_inv = ops.normalize_rows(q, 1e-8, want_y=False)
        logits, cand_inv = ops.logits_dense(q, cand, q_inv, cosine=self.COSINE)
        losses, stats, dl = ops.rowloss(logits, cand.size(1), cfg, mode, tgt, grad_kind,
                                        want_stats=want_stats)
        dq = None
        if dl is not None:
            dq = ops.dq_dense(dl, q, cand, cosine=self.COSINE, q_inv=q_inv, cand_inv=cand_inv)
            if cand_grad:
                self._dcand = ops.dcand_dense(dl, logits, q, cand, cosine=self.COSINE, q_inv=q_inv, cand_inv=cand_inv)
        return losses, stats, dq

    def forward(self, query_embed, candidate_embed, target=None):
        """Summed loss over the batch (xfmr_rec/losses.py:128-155)."""
        losses, _, dq = self._evaluate(query_embed, candidate_embed, target, want_dcand=True)
        loss = losses[N.LOSS_KIND[type(self).__name__]]
        if loss.dtype != torch.float32:
            loss = loss.to(torch.float32)
        dcand, self._dcand = getattr(self, "_dcand", None), None
        if dcand is not None:
            return _AttachGradBoth.apply(query_embed, candidate_embed, loss, dq, dcand)
        if dq is not None:
            src = getattr(query_embed, "_xr_src", None)
            if src is not None and src[0].requires_grad and dq.size(1) % 8 == 0:
                return _AttachGradScatter.apply(src[0], loss, dq, src[1])
            return _AttachGrad.apply(query_embed, loss, dq)
        return loss


def _merge_stats(acc, s):
    if s is None:
        return acc
    if acc is None:
        return s.clone()
    out = acc + s  # sums: slots 0,1,2,3,6,7,8
    out[4] = torch.minimum(acc[4], s[4])
    out[5] = torch.maximum(acc[5], s[5])
    out[9] = torch.minimum(acc[9], s[9])
    out[10] = torch.maximum(acc[10], s[10])
    out[11] = s[11]
    return out


def stats_dict(s: list[float]) -> dict[str, float]:
    """The LogitsStatistics log dict (losses.py:392-404) from the device statistics block: keys and
    empty-set behaviour as the reference (unbiased std; a single element gives nan)."""
    rows = s[1]
    out = {"logits/neg/density": (s[0] / rows) if rows > 0 else float("nan")}
    for key, (cnt, tot, sq, mn, mx) in {
        "pos": (rows, s[2], s[3], s[4], s[5]),
        "neg": (s[6], s[7], s[8], s[9], s[10]),
    }.items():
        if cnt > 0:
            mean = tot / cnt
            var = (sq - cnt * mean * mean) / (cnt - 1) if cnt > 1 else float("nan")
            out[f"logits/{key}/mean"] = mean
            out[f"logits/{key}/std"] = max(var, 0.0) ** 0.5 if var == var else float("nan")
            out[f"logits/{key}/min"] = mn
            out[f"logits/{key}/max"] = mx
    return out


class LogitsStatistics(EmbedLoss):
    """Monitoring statistics over the dot-product logits (xfmr_rec/losses.py:375-405).

    One device pass and ONE device->host copy instead of the reference's nine ``.item()``
    syncs.  Keys and empty-set behaviour follow losses.py:392-404 (std is unbiased; a single
    element gives nan).
    """

    def forward(self, query_embed, candidate_embed, target=None) -> dict[str, float]:
        with torch.no_grad():
            _, stats, _ = self._evaluate(query_embed, candidate_embed, target, want_stats=True)
        return stats_dict(stats.tolist())  # the single host sync


class AlignmentLoss(EmbedLoss):
    """sum_i (1 - cos(q_i, pos_i)) — xfmr_rec/losses.py:408-426."""

    COSINE = True


class AlignmentContrastiveLoss(EmbedLoss):
    """Alignment + margin contrastive on cosine logits (CCL) — xfmr_rec/losses.py:429-447."""

    COSINE = True


class ContrastiveLoss(EmbedLoss):
    """mean_j relu(cos_ij - 1 + margin) over valid negatives — xfmr_rec/losses.py:450-469."""

    COSINE = True


class InfoNCELoss(EmbedLoss):
    """Masked softmax cross-entropy (sampled softmax) — xfmr_rec/losses.py:472-488."""


class NCELoss(EmbedLoss):
    """Binary NCE — xfmr_rec/losses.py:491-511."""


class PairwiseHingeLoss(EmbedLoss):
    """relu(l_neg - (1-margin) l_pos) — xfmr_rec/losses.py:514-527."""


class PairwiseLogisticLoss(EmbedLoss):
    """softplus(l_neg - (1-margin) l_pos); BPR at margin 0 — xfmr_rec/losses.py:530-543."""


LOSS_CLASSES: list[type[EmbedLoss]] = [
    AlignmentLoss,
    AlignmentContrastiveLoss,
    ContrastiveLoss,
    InfoNCELoss,
    NCELoss,
    PairwiseHingeLoss,
    PairwiseLogisticLoss,
]

LossType = Literal[
    "AlignmentLoss",
    "AlignmentContrastiveLoss",
    "ContrastiveLoss",
    "InfoNCELoss",
    "NCELoss",
    "PairwiseHingeLoss",
    "PairwiseLogisticLoss",
]


_MON_TRAIN_KINDS = ("InfoNCELoss", "NCELoss", "PairwiseHingeLoss", "PairwiseLogisticLoss")


def _one_pass(config, query_embed, candidate_embed, target, train_loss):
    """Train loss (a dot-family kind) with its autograd edge AND both families' losses + statistics from ONE tensor-core
    pass (xr_fused_pool_loss_mon) when the batch qualifies: pool candidates on the bf16 tensor-core path for
    both logit families, gradient wanted, no hard-negative mining, scale > 0.  None otherwise."""
    if not (train_loss in _MON_TRAIN_KINDS and target is None and isinstance(candidate_embed, PoolCandidates)
            and torch.is_grad_enabled() and query_embed.requires_grad and query_embed.is_cuda):
        return None
    if config.num_hard_negatives or config.scale <= 0 or config.target_position != "first":
        return None
    dot, cos = InfoNCELoss(config), AlignmentContrastiveLoss(config)
    cdt, logits_bf16 = dot._compute_dtype(query_embed)
    if cdt != torch.bfloat16 or cos._compute_dtype(query_embed)[0] != torch.bfloat16:
        return None
    q, pos, neg = (t.detach().to(cdt).contiguous() for t in (query_embed, candidate_embed.pos, candidate_embed.neg))
    if q.size(0) == 0 or neg.size(0) == 0 or not ops.fused_pool_supported(q, neg):
        return None
    dot.check_embeds(query_embed, candidate_embed)
    cfg = ops.make_cfg(config, logits_bf16=logits_bf16)
    loss, dq, l_dot, l_cos, stats = ops.fused_pool_loss_mon(q, pos, neg, N.LOSS_KIND[train_loss], cfg)
    loss32 = loss.view(torch.float32)[2]
    src = getattr(query_embed, "_xr_src", None)
    if src is not None and src[0].requires_grad and dq.size(1) % 8 == 0:
        attached = _AttachGradScatter.apply(src[0], loss32, dq, src[1])
    else:
        attached = _AttachGrad.apply(query_embed, loss32, dq)
    return attached, l_dot, l_cos, stats


def evaluate_all(config, query_embed, candidate_embed, target=None, *, train_loss="InfoNCELoss"):
    """What ``RecommenderLightningModule.compute_losses`` asks for each step
    (trainer.py:250-263): LogitsStatistics + every loss in LOSS_CLASSES, with the autograd edge
    on ``train_loss`` only.  The two logit families (dot, cosine) are each evaluated ONCE
    instead of the reference's eight separate logit computations.
    Returns (dict loss/<Name> -> 0-dim tensor, stats dict)."""
    out: dict[str, torch.Tensor] = {}
    one = _one_pass(config, query_embed, candidate_embed, target, train_loss)
    if one is not None:
        attached, l_dot, l_cos, stats = one
        for cls in LOSS_CLASSES:
            src = l_cos if cls.COSINE else l_dot
            out[f"loss/{cls.__name__}"] = src[N.LOSS_KIND[cls.__name__]].to(torch.float32)
        out[f"loss/{train_loss}"] = attached
        return out, stats_dict(stats.tolist())
    dot = InfoNCELoss(config)
    cos = AlignmentContrastiveLoss(config)
    with torch.no_grad():
        l_dot, stats, _ = dot._evaluate(query_embed, candidate_embed, target, want_stats=True)
        l_cos, _, _ = cos._evaluate(query_embed, candidate_embed, target, want_stats=True)
    for cls in LOSS_CLASSES:
        src = l_cos if cls.COSINE else l_dot
        out[f"loss/{cls.__name__}"] = src[N.LOSS_KIND[cls.__name__]].to(torch.float32)
    if torch.is_grad_enabled() and query_embed.requires_grad:
        train_cls = {c.__name__: c for c in LOSS_CLASSES}[train_loss]
        out[f"loss/{train_loss}"] = train_cls(config)(query_embed, candidate_embed, target)
    return out, stats_dict(stats.tolist())


def compute_losses(config, embeds: dict, *, train_loss: str = "InfoNCELoss") -> dict:
    """``RecommenderLightningModule.compute_losses`` (xfmr_rec/trainer.py:213-264) from the output of
    ``compute_embeds``: the SAME 30 keys in the same order — ``loss/<Name>`` and ``loss/<Name>Mean``
    for every class in LOSS_CLASSES (0-dim tensors; ``loss/<train_loss>`` carries the autograd edge),
    the seven ``batch/*`` numbers, the nine ``logits/*`` statistics (floats).  Two tensor-core passes
    + the train loss's fused forward/backward and ONE device->host copy, instead of eight logit
    computations and eleven ``.item()`` syncs."""
    attention_mask = embeds["attention_mask"]
    batch_size, seq_len = attention_mask.size()
    numel = attention_mask.numel()
    counts = torch.stack([attention_mask.count_nonzero(), embeds["positive_mask"].count_nonzero()])
    q, cand = embeds["query_embed"], embeds["candidate_embed"]
    one = _one_pass(config, q, cand, None, train_loss)
    attached = None
    if one is not None:      # ONE tensor-core pass: train loss + gradient + both families + statistics
        attached, l_dot, l_cos, stats = one
    else:
        dot, cos = InfoNCELoss(config), AlignmentContrastiveLoss(config)
        with torch.no_grad():
            l_dot, stats, _ = dot._evaluate(q, cand, None, want_stats=True)
            l_cos, _, _ = cos._evaluate(q, cand, None, want_stats=True)
    host = torch.cat([counts.to(torch.float64), stats]).tolist()     # the single host sync
    attn_non_zero, pos_non_zero = int(host[0]), int(host[1])
    losses: dict = {}
    train_cls = {c.__name__: c for c in LOSS_CLASSES}[train_loss]
    for cls in LOSS_CLASSES:
        key = f"loss/{cls.__name__}"
        if cls is train_cls and attached is not None:
            loss = attached
        elif cls is train_cls and torch.is_grad_enabled() and q.requires_grad:
            loss = cls(config)(q, cand)
        else:
            src = l_cos if cls.COSINE else l_dot
            loss = src[N.LOSS_KIND[cls.__name__]].to(torch.float32)
        losses[key] = loss
        losses[f"{key}Mean"] = loss / (pos_non_zero + 1e-9)
    metrics = {
        "batch/size": batch_size,
        "batch/seq_len": seq_len,
        "batch/numel": numel,
        "batch/attention_non_zero": attn_non_zero,
        "batch/attention_density": attn_non_zero / (numel + 1e-9),
        "batch/positive_non_zero": pos_non_zero,
        "batch/positive_density": pos_non_zero / (attn_non_zero + 1e-9),
    }
    return losses | metrics | stats_dict(host[2:])
