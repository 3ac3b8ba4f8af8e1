"""Sequence encoder: the reference's ``BertModel(is_decoder=True)`` on ``inputs_embeds``
from the frozen item table (xfmr_rec/models.py:51-102, 306-345), forward and backward.

Same parameter names and shapes as the HuggingFace module the reference builds (``embeddings.*``,
``encoder.layer.N.attention.self.query.weight`` ...), so a ``BertModel.state_dict()`` — e.g. the
``auto_model`` inside the SentenceTransformer the reference saves (models.py:258-266) — loads with
``load_state_dict`` and vice versa.  ``forward`` returns what the reference's ``RecommenderModel.forward``
returns: ``token_embeddings`` (B, L, H), ``sentence_embedding`` (pooled) and the ``attention_mask``.

What runs where: everything around the linear layers is a hand-written CUDA kernel in
``csrc/encoder.cu`` — the history gather fused with the position / token-type embeddings and the first
LayerNorm, causal multi-head attention (forward and a deterministic backward), exact GELU, and
``LayerNorm(dense_out + bias + residual)`` forward / backward (which also yields the dense bias gradient and
the bf16 copy the next GEMM reads), and the deterministic column sum behind the other bias gradients.  The
linear layers are plain GEMMs and go through cuBLAS (``torch.nn.functional.linear`` / ``@``).  ``compute_dtype=torch.bfloat16`` reproduces Lightning's
``bf16-mixed`` policy: bf16 GEMMs / attention / GELU, fp32 residual stream, LayerNorm and softmax.

Dropout: ``hidden_dropout_prob`` / ``attention_probs_dropout_prob`` (HF defaults 0.1) are active in training
mode, as in HuggingFace: on the embedding LayerNorm output, on the attention probabilities and on the two dense
outputs of every layer.  The keep masks are counter-based (Philox4x32-10 keyed by a seed and a forward counter
that live in device memory): the backward recomputes them, nothing is stored, CUDA-graph replays draw fresh
masks.  ``.eval()`` (or both probabilities 0) gives the deterministic encoder the parity tests pin.
"""

from __future__ import annotations

import ctypes as C
import math
from typing import Literal

import pydantic
import torch
import torch.nn.functional as F

from . import _native as N
from . import ops


class EncoderConfig(pydantic.BaseModel):
    """The ``ModelConfig`` fields that shape the encoder (models.py:22-48); ``hidden_size`` is the item
    embedding width (384 for all-MiniLM-L6-v2, params.py:11)."""

    hidden_size: int = 384
    num_hidden_layers: int = 1
    num_attention_heads: int = 12
    intermediate_size: int = 48
    max_seq_length: int = 32
    layer_norm_eps: float = 1e-12
    # BertConfig defaults (models.py:92-100 passes neither): active in training mode only, as in HuggingFace
    hidden_dropout_prob: float = 0.1
    attention_probs_dropout_prob: float = 0.1
    pooling_mode: Literal["mean", "max", "cls", "lasttoken"] = "mean"
    is_normalized: bool = False


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _pn(t):
    return None if t is None else ops._p(t)


def _grad_pair(dout, dout_lp):
    """The two upstream gradients of a LayerNorm with an fp32 output and a bf16 copy: either may be absent."""
    dout = None if dout is None else dout.contiguous().float()
    dout_lp = None if dout_lp is None else dout_lp.contiguous().to(torch.bfloat16)
    return dout, dout_lp


class _EmbedLN(torch.autograd.Function):
    """x0 = LayerNorm(table[idx] + position_emb[l] + token_type_emb[0]); mask = any(table[idx] != 0).
    Returns (x0 fp32, x0 rounded to bf16 when ``want_lowp`` else None, mask)."""

    @staticmethod
    def forward(ctx, table, idx, pos_w, type_w, ln_w, ln_b, eps, want_lowp, rng=None, drop_p=0.0, site=0):
        dev = ops._require_cuda(table, idx, pos_w, type_w, ln_w, ln_b)
        b, l = idx.shape
        h = table.size(1)
        assert table.dtype == torch.float32 and l <= pos_w.size(0)
        table, idx = table.contiguous(), idx.contiguous()
        out = torch.empty((b, l, h), dtype=torch.float32, device=dev)
        out_lp = torch.empty((b, l, h), dtype=torch.bfloat16, device=dev) if want_lowp else None
        stats = torch.empty((b * l, 2), dtype=torch.float32, device=dev)
        mask = torch.empty((b, l), dtype=torch.uint8, device=dev)
        with ops._on(dev):
            N.call("xr_enc_embed_ln_fwd", ops._p(table), table.size(0), ops._p(idx), ops._p(pos_w), ops._p(type_w),
                   ops._p(ln_w), ops._p(ln_b), b, l, h, float(eps), ops._p(out), _pn(out_lp), ops._p(stats),
                   ops._p(mask), None, _pn(rng), float(drop_p), int(site), ops._stream())
        ctx.save_for_backward(table, idx, pos_w, type_w, ln_w, stats, rng)
        ctx.drop = (float(drop_p), int(site))
        ctx.mark_non_differentiable(mask)
        return out, out_lp, mask

    @staticmethod
    def backward(ctx, dout, dout_lp, _dmask):
        table, idx, pos_w, type_w, ln_w, stats, rng = ctx.saved_tensors
        dev = idx.device
        b, l = idx.shape
        h = table.size(1)
        dout, dout_lp = _grad_pair(dout, dout_lp)
        if dout is None and dout_lp is None:
            return (None,) * 11
        dpos = torch.zeros_like(pos_w)
        dtype = torch.zeros_like(type_w)
        dg, db = torch.empty_like(ln_w), torch.empty_like(ln_w)
        ws = _ws(N.lib().xr_enc_ln_workspace_bytes(b * l), dev)
        dpos_l = torch.empty((l, h), dtype=torch.float32, device=dev)
        dt0 = torch.empty(h, dtype=torch.float32, device=dev)
        with ops._on(dev):
            N.call("xr_enc_embed_ln_bwd", ops._p(table), table.size(0), ops._p(idx), ops._p(pos_w), ops._p(type_w),
                   ops._p(ln_w), ops._p(stats), _pn(dout), _pn(dout_lp), b, l, h, ops._p(dpos_l), ops._p(dt0),
                   ops._p(dg), ops._p(db), ops._p(ws), _pn(rng), ctx.drop[0], ctx.drop[1], ops._stream())
        dpos[:l] = dpos_l
        dtype[0] = dt0
        return None, None, dpos, dtype, dg, db, None, None, None, None, None


class _AddLN(torch.autograd.Function):
    """LayerNorm(y + bias + residual): BertSelfOutput / BertOutput (dropout p = 0) with the dense layer's bias
    add folded in, so that its gradient (the column sums of the LayerNorm input gradient) comes out of the
    same backward pass.  Returns (out fp32, out rounded to bf16 when ``want_lowp`` else None)."""

    @staticmethod
    def forward(ctx, y, bias, residual, ln_w, ln_b, eps, want_lowp, rng=None, drop_p=0.0, site=0):
        dev = ops._require_cuda(y, bias, residual, ln_w, ln_b)
        y, residual = y.contiguous(), residual.contiguous()
        assert residual.dtype == torch.float32 and bias.dtype == torch.float32
        n_tok, h = y.numel() // y.size(-1), y.size(-1)
        out = torch.empty(residual.shape, dtype=torch.float32, device=dev)
        out_lp = torch.empty(residual.shape, dtype=torch.bfloat16, device=dev) if want_lowp else None
        stats = torch.empty((n_tok, 2), dtype=torch.float32, device=dev)
        with ops._on(dev):
            N.call("xr_enc_add_ln_fwd", ops._p(y), ops._dt(y), ops._p(bias), ops._p(residual), ops._p(ln_w),
                   ops._p(ln_b), n_tok, h, float(eps), ops._p(out), _pn(out_lp), ops._p(stats), _pn(rng),
                   float(drop_p), int(site), ops._stream())
        ctx.save_for_backward(y, bias, residual, ln_w, stats, rng)
        ctx.drop = (float(drop_p), int(site))
        return out, out_lp

    @staticmethod
    def backward(ctx, dout, dout_lp):
        y, bias, residual, ln_w, stats, rng = ctx.saved_tensors
        dev = y.device
        n_tok, h = y.numel() // y.size(-1), y.size(-1)
        dout, dout_lp = _grad_pair(dout, dout_lp)
        if dout is None and dout_lp is None:
            return (None,) * 10
        dres = torch.empty_like(residual)
        dy = torch.empty_like(y)
        dbias, dg, db = torch.empty_like(bias), torch.empty_like(ln_w), torch.empty_like(ln_w)
        ws = _ws(N.lib().xr_enc_ln_workspace_bytes(1), dev)
        with ops._on(dev):
            N.call("xr_enc_add_ln_bwd", ops._p(y), ops._dt(y), ops._p(bias), ops._p(residual), ops._p(ln_w),
                   ops._p(stats), _pn(dout), _pn(dout_lp), n_tok, h, ops._p(dres), ops._p(dy), ops._p(dbias),
                   ops._p(dg), ops._p(db), ops._p(ws), _pn(rng), ctx.drop[0], ctx.drop[1], ops._stream())
        return dy, dbias, dres, dg, db, None, None, None, None, None


class _Linear(torch.autograd.Function):
    """y = x W^T (+ bias) in ``x``'s dtype; the weight (fp32 master) is rounded to that dtype per call, as
    autocast does.  The GEMMs (forward, dX, dW) are cuBLAS; the bias gradient is this library's deterministic
    column sum.  ``bias=None``: the consumer (:class:`_AddLN`) adds the bias and produces its gradient."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        wc = weight.to(x.dtype)
        ctx.save_for_backward(x, wc)
        ctx.has_bias = bias is not None
        ctx.master_dtype = weight.dtype
        return F.linear(x, wc, None if bias is None else bias.to(x.dtype))

    @staticmethod
    def backward(ctx, dy):
        x, wc = ctx.saved_tensors
        dy = dy.contiguous().to(x.dtype)
        dy2, x2 = dy.view(-1, dy.size(-1)), x.reshape(-1, x.size(-1))
        dx = (dy2 @ wc).view(x.shape) if ctx.needs_input_grad[0] else None
        dw = (dy2.t() @ x2).to(ctx.master_dtype)
        db = None
        if ctx.has_bias:
            db = torch.empty(dy2.size(1), dtype=torch.float32, device=dy.device)
            ws = _ws(N.lib().xr_enc_colsum_workspace_bytes(dy2.size(1)), dy.device)
            with ops._on(dy.device):
                N.call("xr_enc_colsum", ops._p(dy2), ops._dt(dy2), dy2.size(0), dy2.size(1), ops._p(db), ops._p(ws),
                       ops._stream())
            db = db.to(ctx.master_dtype)
        return dx, dw, db


class _Gelu(torch.autograd.Function):
    """Exact (erf) GELU, hidden_act = "gelu"."""

    @staticmethod
    def forward(ctx, x):
        dev = ops._require_cuda(x)
        x = x.contiguous()
        out = torch.empty_like(x)
        with ops._on(dev):
            N.call("xr_enc_gelu", ops._p(x), None, x.numel(), ops._dt(x), ops._p(out), ops._stream())
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous().to(x.dtype)
        dx = torch.empty_like(x)
        with ops._on(x.device):
            N.call("xr_enc_gelu", ops._p(x), ops._p(dy), x.numel(), ops._dt(x), ops._p(dx), ops._stream())
        return dx


class _Attention(torch.autograd.Function):
    """Causal + key-padding multi-head self-attention on packed qkv (B, L, 3H), head_dim 32."""

    @staticmethod
    def forward(ctx, qkv, mask, n_heads, rng=None, drop_p=0.0, site=0):
        dev = ops._require_cuda(qkv, mask)
        qkv, mask = qkv.contiguous(), mask.contiguous()
        b, l, h3 = qkv.shape
        hid = h3 // 3
        out = torch.empty((b, l, hid), dtype=qkv.dtype, device=dev)
        lse = torch.empty((b, n_heads, l), dtype=torch.float32, device=dev)
        with ops._on(dev):
            N.call("xr_enc_attention", ops._p(qkv), ops._p(mask), None, None, ops._p(lse), b, l, n_heads,
                   hid // n_heads, ops._dt(qkv), ops._p(out), _pn(rng), float(drop_p), int(site), ops._stream())
        ctx.save_for_backward(qkv, mask, out, lse, rng)
        ctx.n_heads = n_heads
        ctx.drop = (float(drop_p), int(site))
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, mask, out, lse, rng = ctx.saved_tensors
        b, l, h3 = qkv.shape
        hid = h3 // 3
        dout = dout.contiguous().to(qkv.dtype)
        dqkv = torch.empty_like(qkv)
        with ops._on(qkv.device):
            N.call("xr_enc_attention", ops._p(qkv), ops._p(mask), ops._p(out), ops._p(dout), ops._p(lse), b, l,
                   ctx.n_heads, hid // ctx.n_heads, ops._dt(qkv), ops._p(dqkv), _pn(rng), ctx.drop[0], ctx.drop[1],
                   ops._stream())
        return dqkv, None, None, None, None, None


# ---- module tree with HuggingFace BertModel's parameter names -------------------------------------------
class _Embeddings(torch.nn.Module):
    def __init__(self, cfg: EncoderConfig):
        super().__init__()
        self.word_embeddings = torch.nn.Embedding(1, cfg.hidden_size)          # vocab_size = 1 (models.py:39): unused
        self.position_embeddings = torch.nn.Embedding(cfg.max_seq_length, cfg.hidden_size)
        self.token_type_embeddings = torch.nn.Embedding(2, cfg.hidden_size)
        self.LayerNorm = torch.nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)


class _SelfAttention(torch.nn.Module):
    def __init__(self, h: int):
        super().__init__()
        self.query, self.key, self.value = (torch.nn.Linear(h, h) for _ in range(3))


class _DenseLN(torch.nn.Module):
    def __init__(self, n_in: int, h: int, eps: float):
        super().__init__()
        self.dense = torch.nn.Linear(n_in, h)
        self.LayerNorm = torch.nn.LayerNorm(h, eps=eps)


class _AttentionBlock(torch.nn.Module):
    def __init__(self, cfg: EncoderConfig):
        super().__init__()
        setattr(self, "self", _SelfAttention(cfg.hidden_size))
        self.output = _DenseLN(cfg.hidden_size, cfg.hidden_size, cfg.layer_norm_eps)


class _Intermediate(torch.nn.Module):
    def __init__(self, cfg: EncoderConfig):
        super().__init__()
        self.dense = torch.nn.Linear(cfg.hidden_size, cfg.intermediate_size)


class _Layer(torch.nn.Module):
    def __init__(self, cfg: EncoderConfig):
        super().__init__()
        self.attention = _AttentionBlock(cfg)
        self.intermediate = _Intermediate(cfg)
        self.output = _DenseLN(cfg.intermediate_size, cfg.hidden_size, cfg.layer_norm_eps)


class _Encoder(torch.nn.Module):
    def __init__(self, cfg: EncoderConfig):
        super().__init__()
        self.layer = torch.nn.ModuleList([_Layer(cfg) for _ in range(cfg.num_hidden_layers)])


class _Pooler(torch.nn.Module):
    def __init__(self, h: int):
        super().__init__()
        self.dense = torch.nn.Linear(h, h)       # BertPooler: present in the checkpoint, unused by the path


class SeqEncoder(torch.nn.Module):
    """``BertModel(is_decoder=True)`` on item embeddings.  ``forward(item_idx, table)``: ``item_idx`` (B, L)
    int64 history indices (0 = padding, right-padded as ``pad_sequence`` does, data.py:801), ``table`` the
    frozen fp32 item table (N+1, H) with a zero row 0 (models.py:247-253)."""

    def __init__(self, config: EncoderConfig | None = None, *, compute_dtype: torch.dtype = torch.float32,
                 seed: int = 0):
        super().__init__()
        self.config = cfg = config or EncoderConfig()
        assert cfg.hidden_size == 384 and cfg.hidden_size // cfg.num_attention_heads == 32, (
            "the encoder kernels are specialised for hidden size 384 / head_dim 32")
        assert compute_dtype in (torch.float32, torch.bfloat16)
        self.compute_dtype = compute_dtype
        self.embeddings = _Embeddings(cfg)
        self.encoder = _Encoder(cfg)
        self.pooler = _Pooler(cfg.hidden_size)
        # dropout state {seed, forward counter}: device memory, so CUDA-graph replays draw fresh masks
        self.register_buffer("_rng", torch.tensor([int(seed), 0], dtype=torch.int64), persistent=False)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):       # BertPreTrainedModel._init_weights: normal(0, 0.02), zero biases, unit LayerNorm
        if isinstance(m, (torch.nn.Linear, torch.nn.Embedding)):
            m.weight.data.normal_(mean=0.0, std=0.02)
            if isinstance(m, torch.nn.Linear):
                m.bias.data.zero_()
        elif isinstance(m, torch.nn.LayerNorm):
            m.weight.data.fill_(1.0)
            m.bias.data.zero_()

    @property
    def max_seq_length(self) -> int:
        return self.config.max_seq_length

    def encode_tokens(self, item_idx: torch.Tensor, table: torch.Tensor):
        """(token_embeddings (B, L, H) fp32, attention_mask (B, L) uint8) for the last ``max_seq_length``
        positions of ``item_idx`` (models.py:334-337).  With ``compute_dtype=bfloat16`` every LayerNorm also
        emits its output rounded to bf16 (the next GEMM's input) and sums the two upstream gradients in its
        backward: no stand-alone cast or add kernels on the residual stream."""
        cfg, cd = self.config, self.compute_dtype
        if not item_idx.is_cuda:
            raise N.NativeError("SeqEncoder needs CUDA tensors; there is no CPU fallback")
        lowp = cd != torch.float32
        idx = item_idx[:, -cfg.max_seq_length:]
        emb = self.embeddings
        # dropout (training mode only): this forward's {seed, counter} snapshot travels to the backward through
        # the autograd contexts; the module's counter advances on the device
        ph = cfg.hidden_dropout_prob if self.training else 0.0
        pa = cfg.attention_probs_dropout_prob if self.training else 0.0
        rng = None
        if ph > 0 or pa > 0:
            rng = self._rng.clone()
            self._rng[1] += 1
        x, x_lp, mask = _EmbedLN.apply(table, idx, emb.position_embeddings.weight, emb.token_type_embeddings.weight,
                                       emb.LayerNorm.weight, emb.LayerNorm.bias, cfg.layer_norm_eps, lowp,
                                       rng, ph, 0)
        n_layers = len(self.encoder.layer)
        for li, layer in enumerate(self.encoder.layer):
            att, att_out, ffn_out = getattr(layer.attention, "self"), layer.attention.output, layer.output
            w = torch.cat([att.query.weight, att.key.weight, att.value.weight], 0)
            b = torch.cat([att.query.bias, att.key.bias, att.value.bias], 0)
            qkv = _Linear.apply(x_lp if lowp else x, w, b)                       # (B, L, 3H)
            ctx = _Attention.apply(qkv, mask, cfg.num_attention_heads, rng, pa, 4 * li + 1)
            x, x_lp = _AddLN.apply(_Linear.apply(ctx, att_out.dense.weight, None), att_out.dense.bias, x,
                                   att_out.LayerNorm.weight, att_out.LayerNorm.bias, cfg.layer_norm_eps, lowp,
                                   rng, ph, 4 * li + 2)
            inter = _Gelu.apply(_Linear.apply(x_lp if lowp else x, layer.intermediate.dense.weight,
                                              layer.intermediate.dense.bias))
            x, x_lp = _AddLN.apply(_Linear.apply(inter, ffn_out.dense.weight, None), ffn_out.dense.bias, x,
                                   ffn_out.LayerNorm.weight, ffn_out.LayerNorm.bias, cfg.layer_norm_eps,
                                   lowp and li + 1 < n_layers, rng, ph, 4 * li + 3)
        return x, mask

    def pool(self, tokens: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        """sentence_transformers ``Pooling`` (models.py:141-145) + optional ``Normalize``."""
        m = mask.to(tokens.dtype)[..., None]
        mode = self.config.pooling_mode
        if mode == "mean":
            out = (tokens * m).sum(1) / m.sum(1).clamp(min=1e-9)
        elif mode == "max":
            out = tokens.masked_fill(m == 0, -1e9).max(1).values
        elif mode == "cls":
            out = tokens[:, 0]
        else:   # lasttoken: the last position with attention (right padding)
            last = (mask.long().sum(1) - 1).clamp(min=0)
            out = tokens[torch.arange(tokens.size(0), device=tokens.device), last]
        return F.normalize(out, p=2, dim=1) if self.config.is_normalized else out

    def forward(self, item_idx: torch.Tensor, table: torch.Tensor) -> dict[str, torch.Tensor]:
        tokens, mask = self.encode_tokens(item_idx, table)
        return {"token_embeddings": tokens, "sentence_embedding": self.pool(tokens, mask),
                "attention_mask": mask.long()}


def encoder_train_step(encoder: SeqEncoder, step, table: torch.Tensor, history_item_idx, pos_item_idx,
                       neg_item_idx):
    """One training step of the whole model path (trainer.py:288-300): encoder forward -> the sync-free
    scoring-and-loss step (:class:`~xfmr_rec_b200.step.PoolLossStep`: gathers, fused contraction + loss,
    dL/d token_embeddings) -> encoder backward, with the loss step's gradient consumed IN PLACE as the
    upstream gradient of ``token_embeddings`` (no autograd node for the loss, no extra copy).  Returns the
    loss (0-dim tensor); parameter gradients are accumulated in ``.grad``."""
    tokens, _ = encoder.encode_tokens(history_item_idx, table)
    loss, dtok = step(tokens.detach(), history_item_idx[:, -encoder.max_seq_length:], pos_item_idx, neg_item_idx)
    tokens.backward(dtok.view_as(tokens).to(tokens.dtype))
    return loss


class GraphedEncoderStep:
    """:func:`encoder_train_step` on static shapes as ONE CUDA graph: encoder forward, the scoring-and-loss
    step and the encoder backward (about 110 kernel launches for two layers) are captured once and replayed
    per batch, so the step costs its GPU time instead of its Python / launch time.

    ``step`` must be built with ``use_graph=False`` (its launches are recorded into this graph).  After a call
    every parameter's ``.grad`` holds THIS batch's gradient (overwritten, not accumulated: the tensors live
    in the graph's memory pool).  ``optimizer`` (optional, built with ``capturable=True``) is stepped inside
    the graph."""

    def __init__(self, encoder: SeqEncoder, step, table: torch.Tensor, history_len: int, *, optimizer=None,
                 warmup: int = 3) -> None:
        if step.graph is not None:
            raise ValueError("GraphedEncoderStep: build the PoolLossStep with use_graph=False")
        dev = table.device
        self.encoder, self.step, self.table, self.optimizer = encoder, step, table, optimizer
        lmax = encoder.max_seq_length
        if min(history_len, lmax) != step.l:
            raise ValueError(f"the loss step is sized for {step.l} positions, the encoder emits {min(history_len, lmax)}")
        self.hist = torch.zeros((step.b, history_len), dtype=torch.int64, device=dev)
        self.pos = torch.zeros((step.b, step.l), dtype=torch.int64, device=dev)
        self.neg = torch.zeros((step.b, step.l), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):          # warm-up: one-time kernel attributes, cuBLAS workspaces
                for _ in range(max(1, warmup)):
                    encoder.zero_grad(set_to_none=True)
                    self._body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(dev)
            if optimizer is not None:
                self._init_optimizer_state(optimizer)
            encoder.zero_grad(set_to_none=True)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss = self._body()
                if optimizer is not None:
                    optimizer.step()

    @staticmethod
    def _init_optimizer_state(optimizer) -> None:
        """Create the optimizer's state tensors BEFORE the capture (a first ``step()`` inside it would record
        their allocation and zero-fill into the graph: every replay would then reset the moments).  One step
        with lr = weight_decay = 0 on the warm-up gradients leaves the parameters untouched; the state it
        created is then zeroed in place, step counters included."""
        saved = [(g["lr"], g["weight_decay"]) for g in optimizer.param_groups]
        for g in optimizer.param_groups:
            g["lr"] = g["lr"] * 0 if torch.is_tensor(g["lr"]) else 0.0
            g["weight_decay"] = 0.0
        optimizer.step()
        for g, (lr, wd) in zip(optimizer.param_groups, saved):
            g["lr"], g["weight_decay"] = lr, wd
        for st in optimizer.state.values():
            for v in st.values():
                if torch.is_tensor(v):
                    v.zero_()

    def _body(self):
        enc = self.encoder
        tokens, _ = enc.encode_tokens(self.hist, self.table)
        loss, dtok = self.step.enqueue(tokens.detach(), self.hist[:, -enc.max_seq_length:], self.pos, self.neg)
        tokens.backward(dtok.view_as(tokens).to(tokens.dtype))
        return loss

    def __call__(self, history_item_idx, pos_item_idx, neg_item_idx) -> torch.Tensor:
        """Copy one batch (device or pinned host tensors) into the static buffers and replay.  Returns the
        loss (0-dim tensor, a view of a static buffer valid until the next call)."""
        self.hist.copy_(history_item_idx, non_blocking=True)
        self.pos.copy_(pos_item_idx, non_blocking=True)
        self.neg.copy_(neg_item_idx, non_blocking=True)
        self.graph.replay()
        return self.loss
