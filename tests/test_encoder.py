"""Sequence encoder against golden vectors produced by executing the reference's own
encoder class — transformers' BertModel(is_decoder=True) on inputs_embeds, built and called as
xfmr_rec/models.py:92-101, 336-345 do (tests/golden/make_golden_encoder.py)."""

import numpy as np
import pytest
import torch

from oracle import xfmr_oracle as orc

HF_KEYS_LAYER = [
    "attention.self.query.weight", "attention.self.query.bias", "attention.self.key.weight",
    "attention.self.key.bias", "attention.self.value.weight", "attention.self.value.bias",
    "attention.output.dense.weight", "attention.output.dense.bias", "attention.output.LayerNorm.weight",
    "attention.output.LayerNorm.bias", "intermediate.dense.weight", "intermediate.dense.bias",
    "output.dense.weight", "output.dense.bias", "output.LayerNorm.weight", "output.LayerNorm.bias"]
HF_KEYS = ["embeddings.word_embeddings.weight", "embeddings.position_embeddings.weight",
           "embeddings.token_type_embeddings.weight", "embeddings.LayerNorm.weight", "embeddings.LayerNorm.bias",
           "pooler.dense.weight", "pooler.dense.bias"]


def seeded_state_dict(ref_sd: dict, seed: int) -> dict:
    """The SAME fill as tests/golden/make_golden_encoder.py (the fixture stores no weights)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(ref_sd):
        shape = tuple(ref_sd[k].shape)
        if k.endswith("LayerNorm.weight"):
            out[k] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            out[k] = 0.05 * torch.randn(shape, generator=g)
    return out


def test_state_dict_keys_are_huggingface_bert_keys():
    """A BertModel checkpoint (the auto_model inside the SentenceTransformer the reference saves,
    models.py:258-266) must load unchanged: same parameter names and shapes."""
    from xfmr_rec_b200.encoder import EncoderConfig, SeqEncoder

    enc = SeqEncoder(EncoderConfig(num_hidden_layers=2, intermediate_size=128, max_seq_length=24))
    want = set(HF_KEYS) | {f"encoder.layer.{i}.{k}" for i in range(2) for k in HF_KEYS_LAYER}
    assert set(enc.state_dict()) == want
    sd = enc.state_dict()
    assert sd["encoder.layer.1.intermediate.dense.weight"].shape == (128, 384)
    assert sd["embeddings.position_embeddings.weight"].shape == (24, 384)
    assert sd["embeddings.token_type_embeddings.weight"].shape == (2, 384)


def test_encoder_has_no_cpu_fallback():
    from xfmr_rec_b200 import _native
    from xfmr_rec_b200.encoder import SeqEncoder

    with pytest.raises(_native.NativeError):
        SeqEncoder()(torch.zeros((2, 5), dtype=torch.int64), torch.zeros((4, 384)))


def _load(golden_dir, tag):
    from xfmr_rec_b200.encoder import EncoderConfig, SeqEncoder

    z = np.load(golden_dir / f"encoder_{tag}.npz")
    layers, inter, max_pos, seed = (int(v) for v in z["config"])
    enc = SeqEncoder(EncoderConfig(num_hidden_layers=layers, intermediate_size=inter, max_seq_length=max_pos))
    enc.load_state_dict(seeded_state_dict(enc.state_dict(), seed))
    return z, enc.cuda().eval()      # the fixtures come from BertModel.eval(): no dropout


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["default_1layer", "2layer_i128"])
def test_encoder_fp32_matches_the_reference_bert(golden_dir, tag):
    """fp32: token embeddings, pooled sentence embeddings and every parameter gradient against
    BertModel's own outputs (1e-5 relative: north_star's fp32 tolerance; gradients 1e-4 of their scale)."""
    z, enc = _load(golden_dir, tag)
    idx, table = torch.from_numpy(z["idx"]).cuda(), torch.from_numpy(z["table"]).cuda()
    out = enc(idx, table)
    tok = out["token_embeddings"]
    want = z["token_embeddings"]
    valid = (z["idx"] != 0)
    assert np.array_equal(out["attention_mask"].cpu().numpy(), valid.astype(np.int64))
    np.testing.assert_allclose(tok.detach().cpu().numpy()[valid], want[valid], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(out["sentence_embedding"].detach().cpu().numpy(), z["sentence_embedding"],
                               rtol=1e-4, atol=2e-5)
    (tok * torch.from_numpy(z["upstream"]).cuda()).sum().backward()
    grads = dict(enc.named_parameters())
    checked = 0
    for key in z.files:
        if key.startswith("grad/"):
            got, ref = grads[key[5:]].grad.cpu().numpy(), z[key]
            if "position_embeddings" in key:
                got = got[: ref.shape[0]]
        elif key.startswith("gradsub/"):
            full = grads[key[8:]].grad
            assert float(full.norm()) == pytest.approx(float(z["gradnorm/" + key[8:]]), rel=1e-4), key
            got, ref = full[::8, ::8].cpu().numpy(), z[key]
        else:
            continue
        scale = max(np.abs(ref).max(), 1e-6)
        # (+ 5e-6 absolute: the key-bias gradient is exactly zero in exact arithmetic -- a shift of every key
        #  leaves the softmax unchanged -- so both sides hold fp32 rounding noise there)
        assert np.abs(got - ref).max() <= 2e-4 * scale + 5e-6, (key, np.abs(got - ref).max(), scale)
        checked += 1
    assert checked >= 20
    assert enc.pooler.dense.weight.grad is None and enc.embeddings.word_embeddings.weight.grad is None


@pytest.mark.gpu
def test_encoder_bf16_mixed_is_close_to_fp32(golden_dir):
    """compute_dtype=bfloat16 (Lightning's bf16-mixed policy: bf16 GEMMs / attention / GELU, fp32 residual
    stream, LayerNorm and softmax): within bf16 tolerance of the fp32 reference."""
    from xfmr_rec_b200.encoder import SeqEncoder

    z, enc = _load(golden_dir, "2layer_i128")
    enc16 = SeqEncoder(enc.config, compute_dtype=torch.bfloat16).cuda().eval()
    enc16.load_state_dict(enc.state_dict())
    idx, table = torch.from_numpy(z["idx"]).cuda(), torch.from_numpy(z["table"]).cuda()
    tok = enc16(idx, table)["token_embeddings"]
    valid = (z["idx"] != 0)
    got, want = tok.detach().cpu().numpy()[valid], z["token_embeddings"][valid]
    assert np.linalg.norm(got - want) <= 2e-2 * np.linalg.norm(want)
    (tok * torch.from_numpy(z["upstream"]).cuda()).sum().backward()
    for key in z.files:
        # (the key-bias gradient is exactly zero in exact arithmetic: nothing to compare but rounding noise)
        if key.startswith("grad/") and "position" not in key and not key.endswith("self.key.bias"):
            got, ref = dict(enc16.named_parameters())[key[5:]].grad.cpu().numpy(), z[key]
            assert np.linalg.norm(got - ref) <= 5e-2 * np.linalg.norm(ref) + 1e-4, key


@pytest.mark.gpu
def test_encoder_is_deterministic_and_truncates(golden_dir):
    """Backward without atomics: two runs give the same bits.  Histories longer than max_seq_length keep
    their LAST max_seq_length positions (models.py:334-337)."""
    z, enc = _load(golden_dir, "default_1layer")
    idx, table = torch.from_numpy(z["idx"]).cuda(), torch.from_numpy(z["table"]).cuda()
    res = []
    for _ in range(2):
        enc.zero_grad()
        tok = enc(idx, table)["token_embeddings"]
        (tok * torch.from_numpy(z["upstream"]).cuda()).sum().backward()
        res.append([tok.detach().clone()] + [p.grad.clone() for p in enc.parameters() if p.grad is not None])
    for a, b in zip(*res):
        assert torch.equal(a, b)
    long_idx = torch.cat([torch.randint(1, 60, (idx.size(0), 40), device="cuda"), idx], 1)[:, -60:]
    a = enc(long_idx, table)["token_embeddings"]
    b = enc(long_idx[:, -32:], table)["token_embeddings"]
    assert a.shape[1] == 32 and torch.equal(a, b)


@pytest.mark.gpu
def test_encoder_train_step_feeds_the_loss_step_gradient_in_place(golden_dir):
    """encoder forward -> PoolLossStep (sync-free scoring-and-loss step) -> encoder backward with the
    step's dL/d token_embeddings as the upstream gradient: same parameter gradients as the module path
    (compute_embeds + loss module + autograd through the same encoder)."""
    import xfmr_rec_b200 as xr
    from xfmr_rec_b200.encoder import EncoderConfig, SeqEncoder, encoder_train_step

    torch.manual_seed(3)
    b, l, n = 6, 24, 500
    table = torch.randn(n + 1, 384, device="cuda") / 384 ** 0.5
    table[0] = 0
    emb = xr.models.ItemEmbeddings(table, add_padding_row=False).cuda()
    lens = torch.randint(2, l + 1, (b,), device="cuda")
    valid = torch.arange(l, device="cuda")[None] < lens[:, None]
    hist = torch.randint(1, n + 1, (b, l), device="cuda") * valid
    pos = torch.randint(1, n + 1, (b, l), device="cuda") * valid
    neg = torch.randint(1, n + 1, (b, l), device="cuda") * valid
    enc = SeqEncoder(EncoderConfig(num_hidden_layers=2, intermediate_size=96, max_seq_length=l)).cuda().eval()
    loss_fn = xr.InfoNCELoss(xr.LossConfig())
    # (fp32 encoder output; logits rounded to bf16 as under bf16-mixed autocast, like the bf16 module path below)
    step = xr.PoolLossStep(emb, loss_fn, b, l, token_dtype=torch.float32, logits_bf16=True)
    loss = encoder_train_step(enc, step, table, hist, pos, neg)
    got = {k: p.grad.clone() for k, p in enc.named_parameters() if p.grad is not None}
    enc.zero_grad()
    tok = enc(hist, table)["token_embeddings"]
    out = xr.models.compute_embeds(emb, tok, hist, pos, neg, candidate_dtype=torch.bfloat16)
    l2 = loss_fn(out["query_embed"].bfloat16(), out["candidate_embed"])
    l2.backward()
    assert float(loss) == pytest.approx(float(l2), rel=1e-5)
    # the module path hands the encoder a bf16-rounded dL/dquery (the loss module returns the query's dtype),
    # the step an fp32 one: parameter gradients agree to bf16 resolution, norm-wise
    for k, p in enc.named_parameters():
        if p.grad is not None and not k.endswith("self.key.bias"):
            assert float((got[k] - p.grad).norm()) <= 1e-2 * float(p.grad.norm()) + 1e-6, k


@pytest.mark.gpu
@pytest.mark.parametrize("seq_len", [1, 37, 200, 384])
def test_attention_tensor_core_path_matches_the_fp32_kernels(seq_len):
    """bf16 activations run the warp-MMA attention kernels, fp32 activations the scalar ones: same masks
    (causal AND key padding, a fully padded sequence included), same bf16-representable inputs, outputs and
    all three gradients within bf16 rounding of the fp32 kernels' (norm-wise 1e-2; the tensor-core path rounds
    P and dS to bf16 before their second contraction).  Ragged lengths exercise the 16-row padding."""
    from xfmr_rec_b200.encoder import _Attention

    g = torch.Generator(device="cuda").manual_seed(seq_len)
    b, heads = 5, 12
    qkv = torch.randn((b, seq_len, 3 * 384), generator=g, device="cuda").bfloat16()
    mask = torch.ones((b, seq_len), dtype=torch.uint8, device="cuda")
    for i in range(b - 1):                      # left padding of different lengths
        mask[i, : (i * seq_len) // 7] = 0
    mask[b - 1] = 0                             # nothing to attend to: zero output, zero gradient
    up = torch.randn((b, seq_len, 384), generator=g, device="cuda").bfloat16()
    res = {}
    for dt in (torch.float32, torch.bfloat16):
        x = qkv.detach().to(dt).clone().requires_grad_(True)
        out = _Attention.apply(x, mask, heads)
        out.backward(up.to(dt))
        res[dt] = (out.detach().float(), x.grad.float())
    for got, want in zip(res[torch.bfloat16], res[torch.float32]):
        assert torch.isfinite(got).all()
        assert float((got - want).norm()) <= 1e-2 * float(want.norm()) + 1e-6
    assert float(res[torch.bfloat16][0][b - 1].abs().max()) == 0.0
    assert float(res[torch.bfloat16][1][b - 1].abs().max()) == 0.0
    x = qkv.detach().clone().requires_grad_(True)
    out2 = _Attention.apply(x, mask, heads)
    out2.backward(up)
    assert torch.equal(out2.float(), res[torch.bfloat16][0]) and torch.equal(x.grad.float(), res[torch.bfloat16][1])


@pytest.mark.gpu
def test_graphed_encoder_step_equals_the_eager_step(golden_dir):
    """One CUDA graph for encoder forward + scoring-and-loss step + encoder backward: same loss and the same
    parameter gradients (bit for bit: every kernel on the path is deterministic) as the eager
    encoder_train_step on two different batches."""
    import xfmr_rec_b200 as xr
    from xfmr_rec_b200.data import synthetic_batch
    from xfmr_rec_b200.encoder import EncoderConfig, GraphedEncoderStep, SeqEncoder, encoder_train_step

    B, L, n_items = 16, 48, 3000
    torch.manual_seed(0)
    enc = SeqEncoder(EncoderConfig(num_hidden_layers=2, intermediate_size=128, max_seq_length=L),
                     compute_dtype=torch.bfloat16).cuda().eval()
    batches = [synthetic_batch(n_items, B, L, dim=384, seed=s) for s in (1, 2)]
    table = torch.from_numpy(batches[0]["table"]).cuda()
    emb = xr.models.ItemEmbeddings(table, add_padding_row=False).cuda()
    mk = lambda **kw: xr.PoolLossStep(emb, xr.InfoNCELoss(xr.LossConfig()), B, L, token_dtype=torch.float32,
                                      logits_bf16=True, **kw)
    eager_step, graphed = mk(), GraphedEncoderStep(enc, mk(use_graph=False), table, L)
    for b in batches:
        hist, pos, neg = (torch.from_numpy(b[k]).cuda() for k in ("history_item_idx", "pos_item_idx", "neg_item_idx"))
        loss_g = graphed(hist, pos, neg).clone()
        grads_g = {n: p.grad.clone() for n, p in enc.named_parameters() if p.grad is not None}
        enc.zero_grad(set_to_none=False)     # the .grad tensors belong to the graph: zero them in place
        loss_e = encoder_train_step(enc, eager_step, table, hist, pos, neg).clone()
        grads_e = {n: p.grad.clone() for n, p in enc.named_parameters() if p.grad is not None}
        assert float(loss_g) == float(loss_e)
        assert grads_g.keys() == grads_e.keys() and len(grads_g) > 30
        for n in grads_g:
            assert torch.equal(grads_g[n], grads_e[n]), n


@pytest.mark.gpu
def test_hidden_dropout_masks_are_seeded_scaled_and_fresh_per_forward(golden_dir):
    """Training mode: dropout(LayerNorm(...)) on the embeddings (BertEmbeddings) keeps a value with
    probability 1 - p_eff and scales it by 1 / (1 - p_eff) (p_eff = round(65536 p) / 65536); the mask is a
    function of (seed, forward counter): two modules with the same seed agree bit for bit, consecutive
    forwards of one module differ; eval mode is the deterministic encoder."""
    from xfmr_rec_b200.encoder import EncoderConfig, SeqEncoder

    z, ref = _load(golden_dir, "default_1layer")
    idx, table = torch.from_numpy(z["idx"]).cuda(), torch.from_numpy(z["table"]).cuda()
    cfg = ref.config.model_copy(update={"hidden_dropout_prob": 0.25, "attention_probs_dropout_prob": 0.0,
                                        "num_hidden_layers": 0})
    mk = lambda seed: SeqEncoder(cfg, seed=seed).cuda()
    a, b, c = mk(7), mk(7), mk(8)
    for m in (a, b, c):
        m.load_state_dict(ref.state_dict(), strict=False)
    base = a.eval()(idx, table)["token_embeddings"]
    a.train()
    out_a1, out_b1, out_c1 = (m(idx, table)["token_embeddings"] for m in (a, b, c))
    out_a2 = a(idx, table)["token_embeddings"]
    assert torch.equal(out_a1, out_b1) and not torch.equal(out_a1, out_c1) and not torch.equal(out_a1, out_a2)
    thr = round(0.25 * 65536)
    scale = 65536.0 / (65536 - thr)
    kept = out_a1 != 0
    assert torch.allclose(out_a1[kept], (base * scale)[kept], rtol=1e-6, atol=0)
    n = base.numel()
    frac = float((~kept & (base != 0)).sum()) / n
    assert abs(frac - thr / 65536) < 4 * (0.25 * 0.75 / n) ** 0.5
    # backward uses the forward's mask: the gradient w.r.t. the LayerNorm weight of a dropped-out network equals
    # autograd through (base-graph output) * mask * scale
    a.zero_grad()
    out = a(idx, table)["token_embeddings"]
    up = torch.randn_like(out)
    (out * up).sum().backward()
    g_drop = a.embeddings.LayerNorm.bias.grad.clone()
    want = (up * (out != 0) * scale).sum((0, 1))        # d out / d beta = mask * scale
    assert torch.allclose(g_drop, want, rtol=1e-4, atol=1e-4)


def _attention_reference(qkv, keymask, keep, scale_drop, up):
    """torch fp32 autograd of causal + key-padding attention with a FIXED dropout mask on the probabilities."""
    b, l, h3 = qkv.shape
    hid = h3 // 3
    x = qkv.detach().float().clone().requires_grad_(True)
    q, k, v = (t.view(b, l, 12, 32).transpose(1, 2) for t in x.split(hid, dim=-1))
    s = q @ k.transpose(-1, -2) / 32 ** 0.5
    allow = torch.tril(torch.ones(l, l, dtype=torch.bool, device=x.device))[None, None] & keymask.bool()[:, None, None, :]
    p = torch.softmax(s.masked_fill(~allow, float("-inf")), dim=-1)
    p = torch.nan_to_num(p, nan=0.0)
    out = ((p * keep * scale_drop) @ v).transpose(1, 2).reshape(b, l, hid)
    out.backward(up.float())
    return out.detach(), x.grad


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_attention_dropout_forward_and_backward_share_one_mask(dtype):
    """Dropout on the attention probabilities (BertSelfAttention): with V = one-hot rows the output IS the
    dropped probability matrix, which gives the mask the kernel drew; the kernel's gradients (dQ pass and
    dK / dV pass recompute the mask in two different fragment layouts) must equal torch autograd with that
    mask held fixed.  Also: kept fraction ~ 1 - p_eff (8-bit draws), scale 1 / (1 - p_eff), fresh mask when the
    counter advances, same mask for the same {seed, counter}."""
    from xfmr_rec_b200.encoder import _Attention

    b, l, heads, p_drop = 3, 32, 12, 0.3
    g = torch.Generator(device="cuda").manual_seed(5)
    qkv = torch.randn((b, l, 3 * 384), generator=g, device="cuda")
    eye = torch.eye(32, device="cuda")[:l]                       # V_j = e_j for every head
    qkv[:, :, 768:] = eye.repeat(1, heads)[None]
    qkv = qkv.to(dtype)
    mask = torch.ones((b, l), dtype=torch.uint8, device="cuda")
    mask[1, :5] = 0
    rng = torch.tensor([11, 3], dtype=torch.int64, device="cuda")
    thr = round(p_drop * 256)
    scale = 256.0 / (256 - thr)
    with torch.no_grad():
        p_eval = _Attention.apply(qkv, mask, heads).float().view(b, l, heads, 32).transpose(1, 2)
        p_drop1 = _Attention.apply(qkv, mask, heads, rng, p_drop, 1).float().view(b, l, heads, 32).transpose(1, 2)
        again = _Attention.apply(qkv, mask, heads, rng.clone(), p_drop, 1).float().view(b, l, heads, 32).transpose(1, 2)
        other = _Attention.apply(qkv, mask, heads, rng + torch.tensor([0, 1], device="cuda"), p_drop, 1).float()
    assert torch.equal(p_drop1, again) and not torch.equal(p_drop1.transpose(1, 2).reshape(b, l, -1), other)
    live = p_eval > 1e-3                                          # entries whose fate is visible
    keep = (p_drop1 != 0)
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-5
    assert torch.allclose(p_drop1[live & keep], (p_eval * scale)[live & keep], rtol=tol, atol=tol)
    n_live = int(live.sum())
    frac = float((live & ~keep).sum()) / n_live
    assert abs(frac - thr / 256) < 4 * (0.3 * 0.7 / n_live) ** 0.5, (frac, thr / 256)
    # gradients with the drawn mask held fixed (entries the probe cannot see are ~0 and do not matter)
    x = qkv.clone().requires_grad_(True)
    up = torch.randn((b, l, 384), generator=g, device="cuda").to(dtype)
    out = _Attention.apply(x, mask, heads, rng, p_drop, 1)
    out.backward(up)
    keep_full = keep | ~live
    want_out, want_grad = _attention_reference(qkv, mask, keep_full.float(), scale, up)
    assert float((out.float() - want_out).norm()) <= (2e-2 if dtype == torch.bfloat16 else 2e-3) * float(want_out.norm())
    err = float((x.grad.float() - want_grad).norm()) / float(want_grad.norm())
    assert err <= (3e-2 if dtype == torch.bfloat16 else 5e-3), err


@pytest.mark.gpu
def test_encoder_trains_with_dropout_and_graph_replays_draw_fresh_masks():
    """Default config (HF's 0.1 / 0.1) in training mode: finite loss and gradients, and two replays of the
    one-graph train step on the SAME batch give different losses (the forward counter lives on the device)."""
    import xfmr_rec_b200 as xr
    from xfmr_rec_b200.data import synthetic_batch
    from xfmr_rec_b200.encoder import EncoderConfig, GraphedEncoderStep, SeqEncoder

    B, L = 8, 40
    b = synthetic_batch(2000, B, L, dim=384, seed=3)
    table = torch.from_numpy(b["table"]).cuda()
    torch.manual_seed(0)
    enc = SeqEncoder(EncoderConfig(num_hidden_layers=2, intermediate_size=128, max_seq_length=L),
                     compute_dtype=torch.bfloat16).cuda().train()
    emb = xr.models.ItemEmbeddings(table, add_padding_row=False).cuda()
    step = xr.PoolLossStep(emb, xr.InfoNCELoss(xr.LossConfig()), B, L, token_dtype=torch.float32, logits_bf16=True,
                           use_graph=False)
    graphed = GraphedEncoderStep(enc, step, table, L)
    hist, pos, neg = (torch.from_numpy(b[k]).cuda() for k in ("history_item_idx", "pos_item_idx", "neg_item_idx"))
    l1 = float(graphed(hist, pos, neg))
    g1 = enc.encoder.layer[0].output.dense.weight.grad.clone()
    l2 = float(graphed(hist, pos, neg))
    g2 = enc.encoder.layer[0].output.dense.weight.grad.clone()
    assert np.isfinite([l1, l2]).all() and l1 != l2
    assert bool(torch.isfinite(g1).all()) and not torch.equal(g1, g2)
    enc.eval()
    with torch.no_grad():
        e1 = enc(hist, table)["token_embeddings"]
        e2 = enc(hist, table)["token_embeddings"]
    assert torch.equal(e1, e2)


@pytest.mark.gpu
def test_training_loop_reduces_the_loss():
    """End to end as INTEGRATION.md 6 / 9 wire it: SeqEncoder (bf16-mixed, dropout on) -> PoolLossStep ->
    encoder backward from the step's dtok -> AdamW.  Forty steps on one small batch must cut the InfoNCE loss
    substantially (the encoder memorises the batch), with finite parameters throughout."""
    import xfmr_rec_b200 as xr
    from xfmr_rec_b200.data import synthetic_batch
    from xfmr_rec_b200.encoder import EncoderConfig, SeqEncoder, encoder_train_step

    B, L = 16, 32
    b = synthetic_batch(400, B, L, dim=384, seed=9)
    table = torch.from_numpy(b["table"]).cuda()
    hist, pos, neg = (torch.from_numpy(b[k]).cuda() for k in ("history_item_idx", "pos_item_idx", "neg_item_idx"))
    torch.manual_seed(1)
    enc = SeqEncoder(EncoderConfig(num_hidden_layers=2, intermediate_size=256, max_seq_length=L),
                     compute_dtype=torch.bfloat16, seed=5).cuda().train()
    emb = xr.models.ItemEmbeddings(table, add_padding_row=False).cuda()
    step = xr.PoolLossStep(emb, xr.InfoNCELoss(xr.LossConfig()), B, L, token_dtype=torch.float32, logits_bf16=True)
    opt = torch.optim.AdamW([p for p in enc.parameters() if p.requires_grad], lr=2e-3, weight_decay=0.0)
    losses = []
    for _ in range(40):
        opt.zero_grad(set_to_none=True)
        losses.append(float(encoder_train_step(enc, step, table, hist, pos, neg)))
        opt.step()
    assert all(np.isfinite(losses)), losses
    assert np.mean(losses[-5:]) < 0.6 * np.mean(losses[:3]), (losses[:3], losses[-5:])
    assert all(bool(torch.isfinite(p).all()) for p in enc.parameters())


@pytest.mark.gpu
def test_evaluate_users_is_encode_then_search_then_metrics(golden_dir):
    """SURVEY 8f ranks 1 + 3 together: evaluate_users = model.encode (sequence encoder, pooled, eval mode) ->
    exact search with the user's whole history excluded -> the seven metrics; equal to the same steps done by
    hand, and the recommended rows equal the oracle's exact search on the encoder's embeddings."""
    import xfmr_rec_b200 as xr

    z, enc = _load(golden_dir, "2layer_i128")
    enc.train()                                              # evaluate_users must switch dropout off itself
    idx_t, table = torch.from_numpy(z["idx"]).cuda(), torch.from_numpy(z["table"]).cuda()
    n_items = table.size(0) - 1
    index = xr.index.ExactIndex(xr.index.ExactIndexConfig(index_metric="cosine", dtype="fp32")).set_catalog(table[1:])
    rng = np.random.default_rng(2)
    targets = [list(map(int, rng.integers(0, n_items, size=int(rng.integers(0, 5))))) for _ in range(idx_t.size(0))]
    means, per_user, valid, rec = xr.evaluate.evaluate_users(enc, table, index, idx_t, targets, 10)
    assert enc.training
    q = enc.eval()(idx_t, table)["sentence_embedding"].detach()
    hist = [[int(x) - 1 for x in row if x != 0] for row in z["idx"]]
    means2, per_user2, valid2, rec2 = xr.evaluate.evaluate_batch(index, q, hist, targets, 10)
    assert torch.equal(rec, rec2) and torch.equal(per_user, per_user2) and torch.equal(valid, valid2)
    _, want = orc.exact_search(q.cpu().numpy(), z["table"][1:], 10, hist, metric="cosine")
    assert np.array_equal(rec.cpu().numpy(), want)
    for r in range(idx_t.size(0)):
        assert not set(rec[r].tolist()) & set(hist[r])


@pytest.mark.gpu
def test_whole_train_step_with_optimizer_as_one_cuda_graph():
    """GraphedEncoderStep(optimizer=AdamW(capturable=True)): encoder forward, scoring-and-loss step, encoder
    backward AND the optimizer update are ONE CUDA graph; replaying it on a batch trains the encoder (the loss
    falls), with fresh dropout masks per replay."""
    import xfmr_rec_b200 as xr
    from xfmr_rec_b200.data import synthetic_batch
    from xfmr_rec_b200.encoder import EncoderConfig, GraphedEncoderStep, SeqEncoder

    B, L = 16, 32
    b = synthetic_batch(400, B, L, dim=384, seed=9)
    table = torch.from_numpy(b["table"]).cuda()
    hist, pos, neg = (torch.from_numpy(b[k]).cuda() for k in ("history_item_idx", "pos_item_idx", "neg_item_idx"))
    torch.manual_seed(1)
    enc = SeqEncoder(EncoderConfig(num_hidden_layers=2, intermediate_size=256, max_seq_length=L),
                     compute_dtype=torch.bfloat16, seed=5).cuda().train()
    trained = [p for n, p in enc.named_parameters() if not n.startswith(("pooler", "embeddings.word"))]
    opt = torch.optim.AdamW(trained, lr=2e-3, weight_decay=0.0, capturable=True)
    emb = xr.models.ItemEmbeddings(table, add_padding_row=False).cuda()
    step = xr.PoolLossStep(emb, xr.InfoNCELoss(xr.LossConfig()), B, L, token_dtype=torch.float32, logits_bf16=True,
                           use_graph=False)
    before = [p.detach().clone() for p in trained]
    graphed = GraphedEncoderStep(enc, step, table, L, optimizer=opt)
    assert all(torch.equal(a, b) for a, b in zip(before, trained))      # building the graph trains nothing
    losses = [float(graphed(hist, pos, neg)) for _ in range(40)]
    assert all(np.isfinite(losses)), losses
    assert np.mean(losses[-5:]) < 0.6 * np.mean(losses[:3]), (losses[:3], losses[-5:])
    # the moments persist across replays (a state created inside the capture would be reset by every replay)
    st = opt.state[trained[0]]
    assert float(st["step"]) == 40 and float(st["exp_avg_sq"].abs().sum()) > 0


@pytest.mark.gpu
def test_gelu_kernels_match_torch_exact_gelu():
    """hidden_act = "gelu" (erf form).  fp32: erff, 1e-6; bf16 activations: the 1.5e-7-accurate polynomial erf,
    outputs and gradients equal torch's exact GELU evaluated in fp32 and rounded to bf16, up to one bf16 ulp
    on a handful of rounding ties."""
    from xfmr_rec_b200.encoder import _Gelu

    g = torch.Generator(device="cuda").manual_seed(0)
    x32 = torch.cat([torch.randn(100_003, generator=g, device="cuda") * 2.5,
                     torch.tensor([0.0, -0.0, 1e-8, -1e-8, 8.0, -8.0, 30.0, -30.0], device="cuda")])
    up32 = torch.randn_like(x32)
    for dt, tol in ((torch.float32, 2e-6), (torch.bfloat16, 8e-3)):
        x = x32.detach().to(dt).clone().requires_grad_(True)
        y = _Gelu.apply(x)
        y.backward(up32.to(dt))
        xr_ = x.detach().float().requires_grad_(True)
        want = torch.nn.functional.gelu(xr_)
        want.backward(up32.to(dt).float())
        assert torch.allclose(y.float(), want.detach().to(dt).float(), rtol=tol, atol=tol * 1e-2 + 1e-7)
        assert torch.allclose(x.grad.float(), xr_.grad.to(dt).float(), rtol=tol, atol=tol)
        assert bool(torch.isfinite(y.float()).all()) and bool(torch.isfinite(x.grad.float()).all())
