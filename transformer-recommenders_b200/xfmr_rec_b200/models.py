"""Candidate construction: the part of ``RecommenderModel`` that is on the hot path.

Mirrors xfmr_rec/models.py:234-259 (frozen item table with a zero padding row) and
models.py:366-419 (``compute_embeds``).  The sequence encoder itself (models.py:22-173,
306-345) is out of scope and drops in unchanged: ``compute_embeds`` here takes its
``token_embeddings`` output.
"""

from __future__ import annotations

import torch

from . import ops
from .losses import PoolCandidates


class _RowGather(torch.autograd.Function):
    """rows = x2d[sel]; backward scatters (the autograd of ``token_embeddings[mask][pos_mask]``,
    models.py:392, 415)."""

    @staticmethod
    def forward(ctx, x2d, sel):
        ctx.save_for_backward(sel)
        ctx.n_rows = x2d.size(0)
        ctx.x_dtype = x2d.dtype
        return ops.gather_rows(x2d, sel)

    @staticmethod
    def backward(ctx, grad):
        (sel,) = ctx.saved_tensors
        return ops.scatter_rows(grad, sel, ctx.n_rows).to(ctx.x_dtype), None


class ItemEmbeddings(torch.nn.Module):
    """Frozen item table, row 0 = padding (``nn.Embedding.from_pretrained(..., freeze=True,
    padding_idx=0)``, models.py:247-253).  Keeps the fp32 master copy, a bf16 copy for the
    tensor-core path and the per-row ``any(row != 0)`` flags that reproduce the attention mask
    of models.py:343 from indices alone."""

    def __init__(self, weights: torch.Tensor, *, add_padding_row: bool = True) -> None:
        super().__init__()
        w = weights.detach().to(torch.float32)
        if add_padding_row:  # models.py:249-250
            w = torch.cat([torch.zeros_like(w[:1]), w])
        self.register_buffer("weight", w.contiguous(), persistent=False)
        self._bf16 = None
        self._rownz = None

    @property
    def num_embeddings(self) -> int:
        return self.weight.size(0)

    @property
    def embedding_dim(self) -> int:
        return self.weight.size(1)

    def weight_bf16(self) -> torch.Tensor:
        if self._bf16 is None or self._bf16.device != self.weight.device:
            self._bf16 = self.weight.to(torch.bfloat16).contiguous()
        return self._bf16

    def rownz(self) -> torch.Tensor:
        if self._rownz is None or self._rownz.device != self.weight.device:
            self._rownz = ops.row_nonzero(self.weight)
        return self._rownz

    def forward(self, idx: torch.Tensor, sel: torch.Tensor | None = None, dtype=None):
        """``self.embeddings(idx)`` of models.py:336-338/400/406 (bit-exact row copies)."""
        dtype = dtype or torch.float32
        table = self.weight_bf16() if dtype == torch.bfloat16 else self.weight
        return ops.gather_rows(table, idx.to(self.weight.device), sel=sel)


def compute_embeds(
    embeddings: ItemEmbeddings,
    token_embeddings: torch.Tensor,
    history_item_idx: torch.Tensor,
    pos_item_idx: torch.Tensor,
    neg_item_idx: torch.Tensor,
    *,
    is_normalized: bool = False,
    dense: bool = False,
    candidate_dtype: torch.dtype = torch.float32,
) -> dict[str, torch.Tensor]:
    """``RecommenderModel.compute_embeds`` (xfmr_rec/models.py:366-419) for a given encoder
    output ``token_embeddings`` (B, L, D).

    Same keys as the reference: ``query_embed`` (M, D, carries autograd to the encoder),
    ``candidate_embed``, ``attention_mask`` (B, L) bool, ``positive_mask`` (M_a,) bool.
    ``candidate_embed`` is a :class:`PoolCandidates` handle unless ``dense=True`` (then the
    reference's (M, 1+M_a, D) tensor is materialised — tests and tiny shapes only).
    """
    dev = embeddings.weight.device
    hist = history_item_idx.to(dev)
    pos = pos_item_idx.to(dev).contiguous().view(-1)
    neg = neg_item_idx.to(dev).contiguous().view(-1)
    b, l = hist.shape
    attention_mask, sel_attn, sel_pos, positive_mask, inv_pos = ops.compact_positions(
        hist, pos, embeddings.rownz(), embeddings.num_embeddings
    )  # models.py:343, 390, 398, 404, 413
    tok2d = token_embeddings.reshape(b * l, token_embeddings.size(-1))
    if tok2d.dtype not in (torch.float32, torch.bfloat16):
        tok2d = tok2d.float()
    tok2d = tok2d.contiguous()
    query_embed = _RowGather.apply(tok2d, sel_pos)  # models.py:392 + :415
    # lets the loss attach its fused backward (scale + cast + scatter in one kernel) straight to
    # the encoder output; lost (on purpose) as soon as the tensor is transformed
    query_embed._xr_src = (tok2d, inv_pos)
    if is_normalized:  # models.py:394-395
        query_embed = torch.nn.functional.normalize(query_embed, dim=-1)
    pos_embed = embeddings(pos, sel=sel_pos, dtype=candidate_dtype)  # models.py:400 + :416
    neg_embed = embeddings(neg, sel=sel_attn, dtype=candidate_dtype)  # models.py:406
    cand = PoolCandidates(pos_embed, neg_embed)
    return {
        "query_embed": query_embed,
        "candidate_embed": cand.dense() if dense else cand,
        "attention_mask": attention_mask,
        "positive_mask": positive_mask,
    }
