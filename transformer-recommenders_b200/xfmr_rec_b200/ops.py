"""Tensor-level wrappers over the C ABI (``include/xfmr_b200.h``).

PyTorch is plumbing here: it owns device memory and the current stream; every function
hands raw device pointers to libxfmr_b200.so.  Inputs must live on a CUDA device — there is
no CPU path.  The main entry points are also registered (at import) as ``torch.library`` custom ops
(``xfmr_b200::gather_rows``, ``::score_loss_fwd_bwd``, ``::pool_loss``, ``::topk``, ``::score_topk``,
``::retrieval_metrics``) with fake implementations and an autograd formula, so they compose with
the dispatcher / autograd / FakeTensor machinery.
"""

from __future__ import annotations

import contextlib
import ctypes as C

import torch

from . import _native as N

_DT = {torch.float32: N.XR_F32, torch.bfloat16: N.XR_BF16}


def _require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise N.NativeError(
                "xfmr_b200 kernels need CUDA tensors (got a tensor on "
                f"{t.device}); there is no CPU fallback"
            )
        dev = dev or t.device
        if t.device != dev:
            raise N.NativeError(f"tensors on different devices: {t.device} vs {dev}")
    return dev


def _p(t: torch.Tensor | None):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    # raw handle of torch's CURRENT stream on the current device (cheap C call; honours
    # torch.cuda.stream(...) contexts)
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


_NULL_CTX = contextlib.nullcontext()


def _on(dev):
    """Device guard that costs nothing when `dev` is already current (the common case)."""
    if dev.index is None or dev.index == torch.cuda.current_device():
        return _NULL_CTX
    return torch.cuda.device(dev)


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise N.NativeError(f"unsupported dtype {t.dtype} (float32 / bfloat16 only)") from None


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def make_cfg(config, *, logits_bf16: bool = False) -> N.XrLossConfig:
    return N.XrLossConfig(
        int(bool(config.mask_false_negatives)),
        int(config.num_hard_negatives),
        float(config.scale),
        float(config.margin),
        int(logits_bf16),
    )


# ---------------------------------------------------------------------------------------------
# family 1: gathers / compaction / normalisation
# ---------------------------------------------------------------------------------------------
def gather_rows(table, idx, sel=None, out_dtype=None, n_out=None, check=False):
    """``table[idx[sel]]`` — nn.Embedding.forward of models.py:336-338/400/406 (bit-exact)."""
    dev = _require_cuda(table, idx, sel)
    table = table.contiguous()
    idx = idx.contiguous()
    assert table.dim() == 2 and idx.dtype == torch.int64
    out_dtype = out_dtype or table.dtype
    lead = idx.shape if sel is None else (sel.numel() if n_out is None else n_out,)
    n = 1
    for s in lead:
        n *= int(s)
    out = torch.empty((n, table.size(1)), dtype=out_dtype, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev) if check else None
    if n == 0:
        return out.view(*lead, table.size(1))
    with _on(dev):
        N.call("xr_gather_rows", _p(table), table.size(0), table.size(1), _dt(table), _p(idx),
               _p(sel), n, _p(out), _DT[out_dtype], _p(err), _stream())
    if check and int(err.item()):
        raise IndexError("index out of range in gather_rows")
    return out.view(*lead, table.size(1))


def scatter_rows(src, sel, n_dst_rows):
    dev = _require_cuda(src, sel)
    src = src.contiguous().float()
    dst = torch.zeros((n_dst_rows, src.size(1)), dtype=torch.float32, device=dev)
    if src.size(0) == 0:
        return dst
    with _on(dev):
        N.call("xr_scatter_rows", _p(src), src.size(0), src.size(1), _p(sel), _p(dst), n_dst_rows,
               _stream())
    return dst


def row_nonzero(table):
    dev = _require_cuda(table)
    table = table.contiguous()
    out = torch.empty(table.size(0), dtype=torch.uint8, device=dev)
    with _on(dev):
        N.call("xr_row_nonzero", _p(table), table.size(0), table.size(1), _dt(table), _p(out),
               _stream())
    return out


def compact_positions(history_idx, pos_idx, rownz=None, n_table_rows=0):
    """models.py:343/390/398/404/413-416 on the index tensors; one host sync for the counts
    (the reference syncs twice for the same numbers, trainer.py:239-240).
    Returns (attention_mask (B,L) bool, sel_attn (M_a,), sel_pos (M,), positive_mask (M_a,) bool,
    inv_pos (B*L,) int64: row of each position in sel_pos or -1)."""
    dev = _require_cuda(history_idx, pos_idx, rownz)
    h = history_idx.contiguous().view(-1)
    p = pos_idx.contiguous().view(-1)
    n = h.numel()
    ibuf = torch.empty(3 * n + 2, dtype=torch.int64, device=dev)   # sel_attn | sel_pos | inv | counts
    bbuf = torch.empty(2 * n, dtype=torch.uint8, device=dev)       # attn | pos_mask
    sel_attn, sel_pos, inv_pos, counts = ibuf[:n], ibuf[n:2 * n], ibuf[2 * n:3 * n], ibuf[3 * n:]
    attn, pos_mask = bbuf[:n], bbuf[n:]
    if n == 0:
        return attn.view(history_idx.shape).view(torch.bool), sel_attn, sel_pos, pos_mask.view(torch.bool), inv_pos
    ws = _ws(N.lib().xr_compact_workspace_bytes(n), dev)
    with _on(dev):
        N.call("xr_compact_positions", _p(h), _p(p), _p(rownz), n_table_rows, n, _p(attn),
               _p(sel_attn), _p(sel_pos), _p(pos_mask), _p(inv_pos), _p(counts), _p(ws), _stream())
    m_a, m = (int(v) for v in counts.tolist())
    return (attn.view(history_idx.shape).view(torch.bool), sel_attn[:m_a], sel_pos[:m],
            pos_mask[:m_a].view(torch.bool), inv_pos)


def scatter_scaled(src, inv_pos, scale, n_dst_rows, out_dtype):
    """dst[p] = inv_pos[p] >= 0 ? cast(src[inv_pos[p]] * scale) : 0 (scale: 0-dim fp32 device tensor
    or None) — the whole backward of the query compaction in one kernel."""
    dev = _require_cuda(src, inv_pos, scale)
    assert src.dtype == torch.float32 and src.is_contiguous()
    dst = torch.empty((n_dst_rows, src.size(1)), dtype=out_dtype, device=dev)
    if scale is not None and scale.dtype != torch.float32:
        scale = scale.float()
    with _on(dev):
        N.call("xr_scatter_scaled", _p(src), _p(inv_pos), _p(scale), n_dst_rows, src.size(1),
               _p(dst), _DT[out_dtype], _stream())
    return dst


def normalize_rows(x, eps=1e-8, out_dtype=None, want_y=True):
    dev = _require_cuda(x)
    x2 = x.contiguous().view(-1, x.size(-1))
    out_dtype = out_dtype or x.dtype
    y = torch.empty(x2.shape, dtype=out_dtype, device=dev) if want_y else None
    inv = torch.empty(x2.size(0), dtype=torch.float32, device=dev)
    if x2.size(0) == 0:
        return (y.view(x.shape) if want_y else None), inv.view(x.shape[:-1])
    with _on(dev):
        N.call("xr_normalize_rows", _p(x2), x2.size(0), x2.size(1), _dt(x2), float(eps), _p(y),
               _DT[out_dtype], _p(inv), _stream())
    return (y.view(x.shape) if want_y else None), inv.view(x.shape[:-1])


# ---------------------------------------------------------------------------------------------
# family 2: logits / rowloss / dq (materialised paths)
# ---------------------------------------------------------------------------------------------
def _ld4(c):
    return (c + 3) // 4 * 4


def logits_pool(q, pos, neg):
    dev = _require_cuda(q, pos, neg)
    m, d = q.shape
    cn = neg.size(0)
    ld = _ld4(cn + 1)
    logits = torch.empty((m, ld), dtype=torch.float32, device=dev)
    with _on(dev):
        N.call("xr_logits_pool", _p(q), _p(pos), _p(neg), m, cn, d, _dt(q), _p(logits), ld,
               _stream())
    return logits  # columns [0,cn) negatives, column cn the positive


def logits_dense(q, cand, q_inv=None, cosine=False, eps=1e-8):
    dev = _require_cuda(q, cand)
    m, c, d = cand.shape
    ld = _ld4(c)
    logits = torch.empty((m, ld), dtype=torch.float32, device=dev)
    cand_inv = torch.empty((m, c), dtype=torch.float32, device=dev) if cosine else None
    with _on(dev):
        N.call("xr_logits_dense", _p(q), _p(cand), m, c, d, _dt(q), _p(q_inv), _p(cand_inv),
               float(eps), _p(logits), ld, _stream())
    return logits, cand_inv


def logits_sampled(q, table, cand_idx, table_inv=None, q_inv=None):
    dev = _require_cuda(q, table, cand_idx)
    m, c = cand_idx.shape
    ld = _ld4(c)
    logits = torch.empty((m, ld), dtype=torch.float32, device=dev)
    with _on(dev):
        N.call("xr_logits_sampled", _p(q), _p(table), table.size(0), _p(cand_idx), m, c, q.size(1),
               _dt(q), _p(table_inv), _p(q_inv), _p(logits), ld, _stream())
    return logits


def rowloss(logits, c, cfg, target_mode, target=None, grad_kind=-1, grad_scale=1.0,
            want_stats=False, check=True):
    """EmbedLoss pipeline on logits (losses.py:211-330 + loss bodies).  Returns
    (losses float64[7] device tensor, stats float64[16] | None, dlogits | None)."""
    dev = _require_cuda(logits, target)
    if logits.stride(1) != 1:
        logits = logits.contiguous()
    m, ld = logits.size(0), logits.stride(0) if logits.size(0) > 1 else logits.size(1)
    losses = torch.empty(N.XR_NUM_LOSSES, dtype=torch.float64, device=dev)
    stats = torch.empty(N.XR_STATS_SLOTS, dtype=torch.float64, device=dev) if want_stats else None
    dlogits = torch.empty((m, ld), dtype=torch.float32, device=dev) if grad_kind >= 0 else None
    err = torch.zeros(1, dtype=torch.int32, device=dev) if (check and target is not None) else None
    ws = _ws(N.lib().xr_rowloss_workspace_bytes(m, c, cfg.num_hard_negatives), dev)
    with _on(dev):
        N.call("xr_rowloss", _p(logits), m, c, ld, target_mode, _p(target), C.byref(cfg), 0x7F,
               grad_kind, float(grad_scale), _p(dlogits), _p(losses), _p(stats), _p(err), _p(ws),
               _stream())
    if err is not None and int(err.item()):
        raise IndexError("target index out of range")
    return losses, stats, dlogits


def dq_pool(dlogits, q, pos, neg, cosine=False, q_inv=None):
    dev = _require_cuda(dlogits, q, pos, neg)
    m, d = q.shape
    dq = torch.empty((m, d), dtype=torch.float32, device=dev)
    with _on(dev):
        N.call("xr_dq_pool", _p(dlogits), dlogits.size(1), _p(q), _p(pos), _p(neg), m, neg.size(0),
               d, _dt(q), int(cosine), _p(q_inv), _p(dq), _stream())
    return dq


def dq_dense(dlogits, q, cand, cosine=False, q_inv=None, cand_inv=None):
    dev = _require_cuda(dlogits, q, cand)
    m, c, d = cand.shape
    dq = torch.empty((m, d), dtype=torch.float32, device=dev)
    with _on(dev):
        N.call("xr_dq_dense", _p(dlogits), dlogits.size(1), _p(q), _p(cand), m, c, d, _dt(q),
               int(cosine), _p(q_inv), _p(cand_inv), _p(dq), _stream())
    return dq


def dcand_dense(dlogits, logits, q, cand, cosine=False, q_inv=None, cand_inv=None):
    """dL/d candidate_embed of a dense (M, C, D) candidate tensor (xr_dcand_dense), fp32."""
    dev = _require_cuda(dlogits, q, cand)
    m, c, d = cand.shape
    out = torch.empty((m, c, d), dtype=torch.float32, device=dev)
    with _on(dev):
        N.call("xr_dcand_dense", _p(dlogits), _p(logits), dlogits.size(1), _p(q), _p(cand), m, c, d, _dt(q),
               int(cosine), _p(q_inv), _p(cand_inv), _p(out), _stream())
    return out


def dq_sampled(dlogits, q, table, cand_idx, table_inv=None, q_inv=None):
    dev = _require_cuda(dlogits, q, table, cand_idx)
    m, c = cand_idx.shape
    dq = torch.empty((m, q.size(1)), dtype=torch.float32, device=dev)
    with _on(dev):
        N.call("xr_dq_sampled", _p(dlogits), dlogits.size(1), _p(q), _p(table), table.size(0),
               _p(cand_idx), m, c, q.size(1), _dt(q), _p(table_inv), _p(q_inv), _p(dq), _stream())
    return dq


def sampled_step(q, table, cand_idx, cfg, table_inv=None, q_inv=None, grad_kind=-1, grad_scale=1.0,
                 want_stats=False):
    """Sampled candidates in one pass (xr_sampled_step): logits + EmbedLoss pipeline + dL/dq with
    the row's logits in shared memory.  Returns (losses f64[7], stats f64[16] | None, dq | None)."""
    dev = _require_cuda(q, table, cand_idx)
    m, c = cand_idx.shape
    losses = torch.empty(N.XR_NUM_LOSSES, dtype=torch.float64, device=dev)
    stats = torch.empty(N.XR_STATS_SLOTS, dtype=torch.float64, device=dev) if want_stats else None
    dq = torch.empty((m, q.size(1)), dtype=torch.float32, device=dev) if grad_kind >= 0 else None
    ws = _ws(N.lib().xr_sampled_step_workspace_bytes(m), dev)
    with _on(dev):
        N.call("xr_sampled_step", _p(q), _p(table), table.size(0), _p(cand_idx), m, c, q.size(1),
               _dt(q), _p(table_inv), _p(q_inv), C.byref(cfg), grad_kind, float(grad_scale), _p(dq),
               _p(losses), _p(stats), _p(ws), _stream())
    return losses, stats, dq


def fused_pool_supported(q, neg) -> bool:
    if not (q.is_cuda and q.dtype == torch.bfloat16 and q.size(1) == 384):
        return False
    major, _ = torch.cuda.get_device_capability(q.device)
    return major == 10 and bool(N.lib().xr_fused_available())


def fused_pool_loss(q, pos, neg, loss_kind, cfg, q_inv=None, grad_scale=1.0, want_grad=True,
                    want_row_loss=False):
    """tcgen05/TMEM fused contraction + loss + dQ (xr_fused_pool_loss).  bf16 inputs."""
    dev = _require_cuda(q, pos, neg, q_inv)
    assert q.dtype == pos.dtype == neg.dtype == torch.bfloat16
    q, pos, neg = q.contiguous(), pos.contiguous(), neg.contiguous()
    m, d = q.shape
    cn = neg.size(0)
    dq = torch.empty((m, d), dtype=torch.float32, device=dev) if want_grad else None
    loss = torch.empty(2, dtype=torch.float64, device=dev)   # [0] f64 sum, [1] carries the f32 copy
    row_loss = torch.empty(m, dtype=torch.float32, device=dev) if want_row_loss else None
    nbytes = N.lib().xr_fused_pool_workspace_bytes(m, cn, d)
    ws = _ws(nbytes, dev)
    with _on(dev):
        N.call("xr_fused_pool_loss", _p(q), _p(pos), _p(neg), m, cn, d, loss_kind, C.byref(cfg),
               _p(q_inv), float(grad_scale), _p(dq), _p(loss), _p(row_loss), _p(ws), ws.numel(),
               _stream())
    return loss, dq, row_loss


def fused_pool_all(q, pos, neg, cfg, cosine):
    """Forward of every loss of one logit family + the LogitsStatistics block in one tcgen05 pass
    (xr_fused_pool_all).  bf16 inputs (pre-normalised rows when ``cosine``).
    Returns (losses f64[7], stats f64[16]) as ``rowloss`` does."""
    dev = _require_cuda(q, pos, neg)
    assert q.dtype == pos.dtype == neg.dtype == torch.bfloat16
    q, pos, neg = q.contiguous(), pos.contiguous(), neg.contiguous()
    m, d = q.shape
    cn = neg.size(0)
    losses = torch.empty(N.XR_NUM_LOSSES, dtype=torch.float64, device=dev)
    stats = torch.empty(N.XR_STATS_SLOTS, dtype=torch.float64, device=dev)
    nbytes = N.lib().xr_fused_pool_all_workspace_bytes(m, cn, d)
    ws = _ws(nbytes, dev)
    with _on(dev):
        N.call("xr_fused_pool_all", _p(q), _p(pos), _p(neg), m, cn, d, int(bool(cosine)),
               C.byref(cfg), _p(losses), _p(stats), _p(ws), ws.numel(), _stream())
    return losses, stats


def fused_pool_loss_mon(q, pos, neg, loss_kind, cfg, want_grad=True, grad_scale=1.0):
    """Train loss (a dot-family kind) + dL/dq AND every loss of both logit families + the LogitsStatistics block from ONE
    tcgen05 pass (xr_fused_pool_loss_mon): what compute_losses (trainer.py:213-264) needs for a batch.
    Returns (loss f64[2] (fp32 copy in the low half of [1]), dq | None, losses_dot f64[7], losses_cos f64[7],
    stats f64[16])."""
    dev = _require_cuda(q, pos, neg)
    assert q.dtype == pos.dtype == neg.dtype == torch.bfloat16
    q, pos, neg = q.contiguous(), pos.contiguous(), neg.contiguous()
    m, d = q.shape
    cn = neg.size(0)
    loss = torch.zeros(2, dtype=torch.float64, device=dev)
    dq = torch.empty((m, d), dtype=torch.float32, device=dev) if want_grad else None
    l_dot = torch.empty(N.XR_NUM_LOSSES, dtype=torch.float64, device=dev)
    l_cos = torch.empty(N.XR_NUM_LOSSES, dtype=torch.float64, device=dev)
    stats = torch.empty(N.XR_STATS_SLOTS, dtype=torch.float64, device=dev)
    ws = _ws(N.lib().xr_fused_pool_loss_mon_workspace_bytes(m, cn, d), dev)
    with _on(dev):
        N.call("xr_fused_pool_loss_mon", _p(q), _p(pos), _p(neg), m, cn, d, int(loss_kind), C.byref(cfg), float(grad_scale),
               _p(dq), _p(loss), _p(l_dot), _p(l_cos), _p(stats), _p(ws), ws.numel(), _stream())
    return loss, dq, l_dot, l_cos, stats


# ---------------------------------------------------------------------------------------------
# family 3: scores / top-k / metrics
# ---------------------------------------------------------------------------------------------
def scores(q, catalog, q_inv=None, cat_inv=None, out=None):
    dev = _require_cuda(q, catalog)
    u, d = q.shape
    n = catalog.size(0)
    ld = _ld4(n)
    if out is None:
        out = torch.empty((u, ld), dtype=torch.float32, device=dev)
    with _on(dev):
        N.call("xr_scores", _p(q), u, _p(catalog), n, d, _dt(q), _p(q_inv), _p(cat_inv), _p(out),
               out.size(1), _stream())
    return out


def _csr(lists, device):
    offs = [0]
    flat: list[int] = []
    for l in lists:
        flat.extend(int(x) for x in l)
        offs.append(len(flat))
    return (torch.tensor(offs, dtype=torch.int64, device=device),
            torch.tensor(flat if flat else [0], dtype=torch.int64, device=device))


def mask_excluded(score_mat, n, exclude_lists, col_offset=0):
    dev = _require_cuda(score_mat)
    offs, ids = exclude_lists if isinstance(exclude_lists, tuple) else _csr(exclude_lists, dev)
    _require_cuda(score_mat, offs, ids)
    with _on(dev):
        N.call("xr_mask_excluded", _p(score_mat), score_mat.size(0), n, score_mat.size(1),
               col_offset, _p(offs), _p(ids), _stream())
    return score_mat


def topk(score_mat, k, n=None, col_offset=0):
    """Exact (score desc, column asc) top-k of each row — stable-sort oracle semantics."""
    dev = _require_cuda(score_mat)
    assert score_mat.dtype == torch.float32 and score_mat.dim() == 2
    score_mat = score_mat if score_mat.stride(1) == 1 else score_mat.contiguous()
    u, ld = score_mat.size(0), score_mat.stride(0)
    n = score_mat.size(1) if n is None else n
    if u <= 1 or ld < n:  # a single row's stride is arbitrary
        score_mat = score_mat.contiguous()
        ld = max(score_mat.size(1), 1)
    out_s = torch.empty((u, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((u, k), dtype=torch.int64, device=dev)
    step = 65535
    for lo in range(0, max(u, 1), step):
        uu = min(step, u - lo)
        if uu <= 0:
            break
        nbytes = N.lib().xr_topk_workspace_bytes(uu, n, k)
        ws = _ws(nbytes, dev)
        with _on(dev):
            N.call("xr_topk", _p(score_mat[lo:]), uu, n, ld, k, col_offset, _p(out_s[lo:]),
                   _p(out_i[lo:]), _p(ws), ws.numel(), _stream())
    return out_s, out_i


def topk_merge(cand_scores, cand_ids, k):
    """Merge (U, G*k) per-shard candidates (global ids) under the same total order."""
    dev = _require_cuda(cand_scores, cand_ids)
    cand_scores, cand_ids = cand_scores.contiguous(), cand_ids.contiguous()
    u, gk = cand_scores.shape
    out_s = torch.empty((u, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((u, k), dtype=torch.int64, device=dev)
    step = 65535   # grid.y limit of the merge kernel, as in topk()
    for lo in range(0, u, step):
        uu = min(step, u - lo)
        ws = _ws(N.lib().xr_topk_merge_workspace_bytes(uu, k), dev)
        with _on(dev):
            N.call("xr_topk_merge", _p(cand_scores[lo:]), _p(cand_ids[lo:]), uu, gk, k, _p(out_s[lo:]),
                   _p(out_i[lo:]), _p(ws), _stream())
    return out_s, out_i


def score_groupmax_supported(q, catalog) -> bool:
    if not (q.is_cuda and q.dtype == torch.bfloat16 and catalog.dtype == torch.bfloat16
            and q.size(1) == 384):
        return False
    major, _ = torch.cuda.get_device_capability(q.device)
    return major == 10 and bool(N.lib().xr_fused_available() & 2)


def score_groupmax(q, catalog, tile_stride=1):
    """tcgen05 scoring without the (U,N) matrix: max score of every 16-row catalog group, natural
    order (column c = rows [16c, 16c+16)); with ``tile_stride`` = s > 1 only every s-th tile of T rows
    (T = 128 for U > 128, else 64) is scored and column (T/16) t + g = rows [T t s + 16 g, +16)."""
    dev = _require_cuda(q, catalog)
    q, catalog = q.contiguous(), catalog.contiguous()
    u, d = q.shape
    n = catalog.size(0)
    ld = int(N.lib().xr_score_groupmax_ld(u, n, tile_stride))
    ld += ld % 2
    gmax = torch.empty((u, ld), dtype=torch.float32, device=dev)
    with _on(dev):
        N.call("xr_score_groupmax", _p(q), u, _p(catalog), n, d, tile_stride, _p(gmax), ld, _stream())
    return gmax


def groups_to_rows(group_ids, n, row_offset=0):
    """(U, kg) group ids -> (cols, ids), both (U, kg*16) int64: gather rows and global ids (-1 = none)."""
    dev = _require_cuda(group_ids)
    group_ids = group_ids.contiguous()
    u, kg = group_ids.shape
    cols = torch.empty((u, kg * 16), dtype=torch.int64, device=dev)
    ids = torch.empty((u, kg * 16), dtype=torch.int64, device=dev)
    with _on(dev):
        N.call("xr_groups_to_rows", _p(group_ids), u, kg, n, row_offset, _p(cols), _p(ids), _stream())
    return cols, ids


def mask_excluded_ids(score_mat, ids, id_lo, id_hi, exclude=None):
    offs, ex = (None, None) if exclude is None else exclude
    dev = _require_cuda(score_mat, ids, offs, ex)
    u, c = ids.shape
    with _on(dev):
        N.call("xr_mask_excluded_ids", _p(score_mat), _p(ids), u, c, score_mat.size(1), id_lo, id_hi,
               _p(offs), _p(ex), _stream())
    return score_mat


class FilterSurvivors:
    """Survivor storage of xr_score_filter: ``n_sub`` sub-buckets of ``cap_b`` slots per query (each filled
    by one lane of the scoring kernel, no atomics) + one overflow list per query."""

    def __init__(self, u, n_sub, cap_b, ovf_cap, device):
        self.u, self.n_sub, self.cap_b, self.ovf_cap = u, n_sub, cap_b, ovf_cap
        self.b_scores = torch.empty((u, n_sub, cap_b), dtype=torch.float32, device=device)
        self.b_rows = torch.empty((u, n_sub, cap_b), dtype=torch.int32, device=device)
        self.b_count = torch.empty((u, n_sub), dtype=torch.int32, device=device)
        self.o_scores = torch.empty((u, ovf_cap), dtype=torch.float32, device=device)
        self.o_rows = torch.empty((u, ovf_cap), dtype=torch.int32, device=device)
        self.o_count = torch.zeros(u, dtype=torch.int32, device=device)

    def counts(self) -> torch.Tensor:
        """survivors per query (sub-bucket counts include what spilled to the overflow list)"""
        return self.b_count.sum(1)

    def lists(self):
        """per query: (scores, rows) numpy arrays of every stored survivor (tests / diagnostics)."""
        bc = self.b_count.cpu().numpy()
        bs, br = self.b_scores.cpu().numpy(), self.b_rows.cpu().numpy()
        oc = self.o_count.cpu().numpy()
        os_, or_ = self.o_scores.cpu().numpy(), self.o_rows.cpu().numpy()
        import numpy as np

        out = []
        for r in range(self.u):
            sc = [bs[r, s, :min(c, self.cap_b)] for s, c in enumerate(bc[r])] + [os_[r, :min(oc[r], self.ovf_cap)]]
            ro = [br[r, s, :min(c, self.cap_b)] for s, c in enumerate(bc[r])] + [or_[r, :min(oc[r], self.ovf_cap)]]
            out.append((np.concatenate(sc), np.concatenate(ro)))
        return out


def score_filter(q, catalog, thresh, expected_survivors=4096, cap_b=None, ovf_cap=8192):
    """Scoring pass with the threshold filter in the epilogue (xr_score_filter): returns the survivor
    storage (:class:`FilterSurvivors`)."""
    dev = _require_cuda(q, catalog, thresh)
    q, catalog = q.contiguous(), catalog.contiguous()
    thresh = thresh.contiguous().float()
    u, n = q.size(0), catalog.size(0)
    n_sub, cb = C.c_int64(), C.c_int64()
    N.call("xr_score_filter_layout", u, n, expected_survivors, C.byref(n_sub), C.byref(cb))
    fs = FilterSurvivors(u, n_sub.value, cap_b or cb.value, ovf_cap, dev)
    with _on(dev):
        N.call("xr_score_filter", _p(q), u, _p(catalog), n, q.size(1), _p(thresh), 1, _p(fs.b_scores),
               _p(fs.b_rows), _p(fs.b_count), fs.n_sub, fs.cap_b, _p(fs.o_scores), _p(fs.o_rows),
               _p(fs.o_count), fs.ovf_cap, _stream())
    return fs


def filter_finalize(fs, n, thresh, k_sel, k, row_offset=0, exclude=None, max_excl=0, flags=None):
    dev = _require_cuda(fs.b_scores, thresh)
    thresh = thresh.contiguous().float()
    offs, ex = (None, None) if exclude is None else exclude
    u = fs.u
    out_s = torch.empty((u, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((u, k), dtype=torch.int64, device=dev)
    if flags is None:
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
    with _on(dev):
        N.call("xr_filter_finalize", u, n, _p(fs.b_scores), _p(fs.b_rows), _p(fs.b_count), fs.n_sub, fs.cap_b,
               _p(fs.o_scores), _p(fs.o_rows), _p(fs.o_count), fs.ovf_cap, _p(thresh), 1, k_sel, k, row_offset,
               _p(offs), _p(ex), max_excl, _p(out_s), _p(out_i), _p(flags), _stream())
    return out_s, out_i, flags


def kth_largest(x, kth, n=None):
    """kth largest of every row of a (U, ld) fp32 matrix (first n columns): xr_kth_largest."""
    dev = _require_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    u = x.size(0)
    n = x.size(1) if n is None else n
    out = torch.empty(u, dtype=torch.float32, device=dev)
    with _on(dev):
        N.call("xr_kth_largest", _p(x), u, n, x.stride(0) if u > 1 else x.size(1), kth, _p(out), _stream())
    return out


SCORE_TOPK_MAX_K = 1024   # k + max_excl the one-call search serves (xr_score_topk)


def score_topk_supported(q, catalog) -> bool:
    return score_groupmax_supported(q, catalog)


def score_topk(q, catalog, k, row_offset=0, exclude=None, max_excl=0, flags=None, out=None, ws=None):
    """The whole local search of one catalog shard in one call (xr_score_topk): sample thresholds ->
    filter in the scoring epilogue -> exact top-k of the survivors.  ``exclude``: CSR pair of GLOBAL
    ids with at most ``max_excl`` per query.  ``flags`` (int32[1], sticky) becomes non-zero when the
    result is not guaranteed exact (survivor list overflow / more exclusions than announced): the
    caller then takes the materialised path.  Returns (scores (U,k), ids (U,k) global, flags)."""
    dev = _require_cuda(q, catalog)
    q, catalog = q.contiguous(), catalog.contiguous()
    u, d = q.shape
    n = catalog.size(0)
    offs, ids = (None, None) if exclude is None else exclude
    if out is None:
        out = (torch.empty((u, k), dtype=torch.float32, device=dev),
               torch.empty((u, k), dtype=torch.int64, device=dev))
    if flags is None:
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
    if ws is None:
        ws = score_topk_workspace(u, n, k, max_excl, dev)
    ptr = ws.data_ptr() + (-ws.data_ptr()) % 256
    with _on(dev):
        N.call("xr_score_topk", _p(q), u, _p(catalog), n, d, k, row_offset, _p(offs), _p(ids), max_excl,
               _p(out[0]), _p(out[1]), _p(flags), C.c_void_p(ptr), ws.numel() - 256, _stream())
    return out[0], out[1], flags


def score_topk_workspace(u, n, k, max_excl, device):
    return _ws(N.lib().xr_score_topk_workspace_bytes(u, n, k, max_excl) + 256, device)


def retrieval_metrics(rec_idx, target_lists, top_k):
    """metrics.py:62-79 batched.  rec_idx (U,k) int64 (-1 = padding); returns ((U,7) fp32, valid)."""
    dev = _require_cuda(rec_idx)
    rec_idx = rec_idx.contiguous()
    u, k = rec_idx.shape
    offs, ids = target_lists if isinstance(target_lists, tuple) else _csr(target_lists, dev)
    out = torch.empty((u, 7), dtype=torch.float32, device=dev)
    valid = torch.empty(u, dtype=torch.uint8, device=dev)
    with _on(dev):
        N.call("xr_retrieval_metrics", _p(rec_idx), u, k, _p(offs), _p(ids), top_k, _p(out),
               _p(valid), _stream())
    return out, valid.bool()


# ---------------------------------------------------------------------------------------------
# torch.library registration (the dispatcher-visible names of SURVEY §8b), done at import:
#   xfmr_b200::gather_rows, ::score_loss_fwd_bwd, ::pool_loss (autograd), ::topk, ::score_topk,
#   ::retrieval_metrics.  Each has a fake (meta) implementation, so the ops trace under FakeTensor /
#   torch.export, and ::pool_loss carries an autograd formula (the gradient is produced by the same
#   kernel launch as the loss; backward only scales it).  CUDA is the only backend registered: on any
#   other device the dispatcher raises, as everywhere in this package there is no fallback.
# ---------------------------------------------------------------------------------------------
def _lossname_to_cfg(mask_fn: bool, scale: float, margin: float, logits_bf16: bool) -> N.XrLossConfig:
    return N.XrLossConfig(int(mask_fn), 0, float(scale), float(margin), int(logits_bf16))


@torch.library.custom_op("xfmr_b200::gather_rows", mutates_args=(), device_types="cuda")
def _op_gather_rows(table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    return gather_rows(table, idx)


@_op_gather_rows.register_fake
def _(table, idx):
    return table.new_empty((*idx.shape, table.size(1)))


@torch.library.custom_op("xfmr_b200::score_loss_fwd_bwd", mutates_args=(), device_types="cuda")
def _op_score_loss_fwd_bwd(q: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor, loss_kind: int,
                           mask_fn: bool, scale: float, margin: float,
                           logits_bf16: bool) -> tuple[torch.Tensor, torch.Tensor]:
    """Fused contraction + loss + dL/dq over a shared negative pool (bf16 operands, D = 384):
    (loss 0-dim fp32, dq (M, D) fp32)."""
    loss, dq, _ = fused_pool_loss(q, pos, neg, loss_kind, _lossname_to_cfg(mask_fn, scale, margin, logits_bf16))
    return loss.view(torch.float32)[2].clone(), dq


@_op_score_loss_fwd_bwd.register_fake
def _(q, pos, neg, loss_kind, mask_fn, scale, margin, logits_bf16):
    return q.new_empty((), dtype=torch.float32), q.new_empty(q.shape, dtype=torch.float32)


@torch.library.custom_op("xfmr_b200::pool_loss", mutates_args=(), device_types="cuda")
def _op_pool_loss(q: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor, loss_kind: int, mask_fn: bool,
                  scale: float, margin: float, logits_bf16: bool) -> tuple[torch.Tensor, torch.Tensor]:
    """Differentiable form: returns (loss, dq); dq is saved for backward (autograd formula below)."""
    return torch.ops.xfmr_b200.score_loss_fwd_bwd(q, pos, neg, loss_kind, mask_fn, scale, margin, logits_bf16)


@_op_pool_loss.register_fake
def _(q, pos, neg, loss_kind, mask_fn, scale, margin, logits_bf16):
    return q.new_empty((), dtype=torch.float32), q.new_empty(q.shape, dtype=torch.float32)


def _pool_loss_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1])
    ctx.q_dtype = inputs[0].dtype


def _pool_loss_backward(ctx, grad_loss, grad_dq):
    (dq,) = ctx.saved_tensors
    return (dq * grad_loss).to(ctx.q_dtype), None, None, None, None, None, None, None


_op_pool_loss.register_autograd(_pool_loss_backward, setup_context=_pool_loss_setup)


@torch.library.custom_op("xfmr_b200::topk", mutates_args=(), device_types="cuda")
def _op_topk(scores: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor]:
    return topk(scores, k)


@_op_topk.register_fake
def _(scores, k):
    return (scores.new_empty((scores.size(0), k), dtype=torch.float32),
            scores.new_empty((scores.size(0), k), dtype=torch.int64))


@torch.library.custom_op("xfmr_b200::score_topk", mutates_args=(), device_types="cuda")
def _op_score_topk(q: torch.Tensor, catalog: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor]:
    """Whole local search of one catalog shard (bf16, D = 384, rows pre-normalised for cosine).  Falls
    back to NOTHING: if the filter path cannot vouch for the result the op raises."""
    s, i, flags = score_topk(q, catalog, k)
    if int(flags.item()):
        raise N.NativeError("xfmr_b200::score_topk: survivor list overflow (massively tied scores); "
                            "use ExactIndex.search_batch, which takes the materialised path")
    return s, i


@_op_score_topk.register_fake
def _(q, catalog, k):
    return (q.new_empty((q.size(0), k), dtype=torch.float32), q.new_empty((q.size(0), k), dtype=torch.int64))


@torch.library.custom_op("xfmr_b200::retrieval_metrics", mutates_args=(), device_types="cuda")
def _op_retrieval_metrics(rec_idx: torch.Tensor, target_offsets: torch.Tensor, target_ids: torch.Tensor,
                          top_k: int) -> tuple[torch.Tensor, torch.Tensor]:
    return retrieval_metrics(rec_idx, (target_offsets, target_ids), top_k)


@_op_retrieval_metrics.register_fake
def _(rec_idx, target_offsets, target_ids, top_k):
    return (rec_idx.new_empty((rec_idx.size(0), 7), dtype=torch.float32),
            rec_idx.new_empty((rec_idx.size(0),), dtype=torch.bool))


CUSTOM_OPS = ("gather_rows", "score_loss_fwd_bwd", "pool_loss", "topk", "score_topk", "retrieval_metrics")
