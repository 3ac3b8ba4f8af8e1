"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden
vectors produced by the reference's own losses.py.  Run on the B200 box: pytest -m gpu."""

import json
import math

import numpy as np
import pytest
import torch

from oracle import xfmr_oracle as orc

pytestmark = pytest.mark.gpu

FP32_REL = 1e-5   # north_star: losses / gradients within 1e-5 relative in fp32
BF16_REL = 2e-3   # ... and 2e-3 in bf16


@pytest.fixture(scope="module")
def xr():
    import xfmr_rec_b200 as pkg

    return pkg


def dev(x, dtype=None):
    t = torch.as_tensor(x).cuda()
    return t.to(dtype) if dtype is not None else t


def load(golden_dir, name):
    z = np.load(golden_dir / f"losses_{name}.npz")
    return z, json.loads(str(z["cfg"]))


def assert_close_grad(got, want, rel):
    scale = max(float(np.abs(want).max()), 1e-6)
    err = float(np.abs(got - want).max())
    assert err <= rel * scale * 4 + 1e-7, (err, scale)


# ------------------------------------------------------------------------------------------
# family 1: gathers / compaction
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,dtype", [(384, torch.float32), (384, torch.bfloat16), (50, torch.float32),
                                        (8, torch.bfloat16), (1024, torch.float32)])
def test_gather_rows_bit_exact(xr, dim, dtype):
    g = torch.Generator().manual_seed(0)
    table = torch.randn(1000, dim, generator=g).to(dtype)
    table[0] = 0
    idx = torch.randint(0, 1000, (37, 11), generator=g)
    out = xr.ops.gather_rows(table.cuda(), idx.cuda())
    assert out.shape == (37, 11, dim)
    assert torch.equal(out.cpu(), table[idx])           # bit-exact (nn.Embedding semantics)
    # empty and single-row inputs
    assert xr.ops.gather_rows(table.cuda(), idx[:0].cuda()).shape == (0, 11, dim)
    sel = torch.tensor([5, 0, 5, 36 * 11], dtype=torch.int64)
    out = xr.ops.gather_rows(table.cuda(), idx.cuda().view(-1), sel=sel.cuda())
    assert torch.equal(out.cpu(), table[idx.view(-1)[sel]])


def test_gather_rows_cast_and_bounds(xr):
    g = torch.Generator().manual_seed(1)
    table = torch.randn(300, 384, generator=g)
    idx = torch.randint(0, 300, (1000,), generator=g)
    out = xr.ops.gather_rows(table.cuda(), idx.cuda(), out_dtype=torch.bfloat16)
    assert torch.equal(out.cpu(), table[idx].bfloat16())  # round-to-nearest-even like torch
    bad = idx.clone()
    bad[3] = 300
    with pytest.raises(IndexError):
        xr.ops.gather_rows(table.cuda(), bad.cuda(), check=True)


def test_gather_large_matches_torch(xr):
    g = torch.Generator().manual_seed(2)
    table = torch.randn(27279, 384, generator=g).bfloat16().cuda()
    idx = torch.randint(0, 27279, (25600,), generator=g).cuda()
    assert torch.equal(xr.ops.gather_rows(table, idx), table[idx])


@pytest.mark.parametrize("batch,seq", [(4, 9), (32, 50), (3, 4097)])
def test_compute_embeds_matches_oracle(xr, batch, seq):
    b = orc.synth_batch(200, batch, seq, dim=64, seed=batch)
    b["table"][17] = 0.0    # a real item with an all-zero row: masked by (embeds != 0).any(-1)
    want = orc.compute_embeds(b["table"], b["token_embeddings"], b["history_item_idx"],
                              b["pos_item_idx"], b["neg_item_idx"], dense=False)
    emb = xr.models.ItemEmbeddings(torch.from_numpy(b["table"]), add_padding_row=False).cuda()
    tok = dev(b["token_embeddings"]).requires_grad_(True)
    out = xr.models.compute_embeds(emb, tok, dev(b["history_item_idx"]), dev(b["pos_item_idx"]),
                                   dev(b["neg_item_idx"]))
    assert np.array_equal(out["attention_mask"].cpu().numpy(), want["attention_mask"])
    assert np.array_equal(out["positive_mask"].cpu().numpy(), want["positive_mask"])
    assert np.array_equal(out["query_embed"].detach().cpu().numpy(), want["query_embed"])
    assert np.array_equal(out["candidate_embed"].pos.cpu().numpy(), want["pos_embed"])
    assert np.array_equal(out["candidate_embed"].neg.cpu().numpy(), want["neg_embed"])
    assert tuple(out["candidate_embed"].size()) == (want["query_embed"].shape[0],
                                                    1 + want["neg_embed"].shape[0], 64)
    # autograd reaches the encoder output through the compaction
    out["query_embed"].sum().backward()
    g = tok.grad.cpu().numpy().reshape(-1, 64)
    am = want["attention_mask"].reshape(-1)
    keep = np.zeros(am.shape, bool)
    keep[np.flatnonzero(am)[want["positive_mask"]]] = True
    assert np.array_equal(g, np.repeat(keep[:, None], 64, 1).astype(np.float32))


@pytest.mark.parametrize("case", ["basic", "normalized", "truncated"])
def test_compute_embeds_matches_reference_outputs(xr, golden_dir, case):
    """Against the outputs of the reference's OWN forward + compute_embeds (models.py:306-345,
    366-419; tests/golden/make_golden_embeds.py): rows bit-exact, masks equal, and the autograd
    of the query selection back to the encoder output."""
    z = np.load(golden_dir / f"embeds_{case}.npz")
    emb = xr.models.ItemEmbeddings(torch.from_numpy(z["table"]), add_padding_row=False).cuda()
    tok = torch.from_numpy(z["tokens"]).cuda().requires_grad_(True)
    args = [torch.from_numpy(z[k]).cuda() for k in ("history_item_idx", "pos_item_idx", "neg_item_idx")]
    norm = bool(z["is_normalized"])
    out = xr.models.compute_embeds(emb, tok, *args, is_normalized=norm, dense=True)
    assert np.array_equal(out["attention_mask"].cpu().numpy(), z["attention_mask"].astype(bool))
    assert np.array_equal(out["positive_mask"].cpu().numpy(), z["positive_mask"].astype(bool))
    assert np.array_equal(out["candidate_embed"].cpu().numpy(), z["candidate_embed"])
    q = out["query_embed"]
    if norm:
        np.testing.assert_allclose(q.detach().cpu().numpy(), z["query_embed"], rtol=1e-6, atol=1e-7)
    else:
        assert np.array_equal(q.detach().cpu().numpy(), z["query_embed"])
    (q * torch.from_numpy(z["upstream"]).cuda()).sum().backward()
    np.testing.assert_allclose(tok.grad.cpu().numpy(), z["dtokens"], rtol=1e-5, atol=1e-7)
    handle = xr.models.compute_embeds(emb, tok.detach(), *args, is_normalized=norm)["candidate_embed"]
    assert np.array_equal(handle.pos.cpu().numpy(), z["candidate_embed"][:, 0])
    if z["candidate_embed"].shape[0]:
        assert np.array_equal(handle.neg.cpu().numpy(), z["candidate_embed"][0, 1:])


def test_compute_embeds_dense_equals_reference_layout(xr):
    b = orc.synth_batch(60, 3, 7, dim=32, seed=5)
    want = orc.compute_embeds(b["table"], b["token_embeddings"], b["history_item_idx"],
                              b["pos_item_idx"], b["neg_item_idx"], dense=True)
    emb = xr.models.ItemEmbeddings(torch.from_numpy(b["table"]), add_padding_row=False).cuda()
    out = xr.models.compute_embeds(emb, dev(b["token_embeddings"]), dev(b["history_item_idx"]),
                                   dev(b["pos_item_idx"]), dev(b["neg_item_idx"]), dense=True)
    assert np.array_equal(out["candidate_embed"].cpu().numpy(), want["candidate_embed"])


# ------------------------------------------------------------------------------------------
# family 2: losses vs the reference's own outputs (golden) — fp32
# ------------------------------------------------------------------------------------------
GOLDEN_FP32 = ["anchor_m64_d384", "pool_default", "pool_nomask", "pool_scale20_margin02",
               "pool_margin0", "pool_hard5", "pool_hard5_nomask", "dense_first", "dense_diagonal",
               "dense_explicit", "dense_c1"]


@pytest.mark.parametrize("case", GOLDEN_FP32)
@pytest.mark.parametrize("form", ["dense", "handle"])
def test_losses_match_reference_fp32(xr, golden_dir, case, form):
    z, cfg_kw = load(golden_dir, case)
    if form == "handle" and "pos" not in z:
        pytest.skip("dense-only case")
    cfg = xr.LossConfig(**cfg_kw)
    target = dev(z["target"]) if "target" in z else None
    if form == "handle":
        cand = xr.PoolCandidates(dev(z["pos"]), dev(z["neg"]))
    elif "cand" in z:
        cand = dev(z["cand"])
    else:
        cand = xr.PoolCandidates(dev(z["pos"]), dev(z["neg"])).dense()
    for cls in xr.LOSS_CLASSES:
        name = cls.__name__
        q = dev(z["query"]).requires_grad_(True)
        loss = cls(cfg)(query_embed=q, candidate_embed=cand, target=target)
        assert loss.dim() == 0 and loss.dtype == torch.float32
        want = float(z[f"loss/{name}"])
        assert float(loss) == pytest.approx(want, rel=FP32_REL, abs=1e-5), (case, name)
        loss.backward()
        assert_close_grad(q.grad.cpu().numpy(), z[f"dq/{name}"], FP32_REL)


@pytest.mark.parametrize("case", ["dcand_default", "dcand_explicit_scaled"])
def test_candidate_gradient_matches_reference(xr, golden_dir, case):
    """The reference's losses are differentiable in BOTH arguments (losses.py:128-155): for the dense
    (M, C, D) candidate tensor of its own API, dL/dcandidate_embed (xr_dcand_dense) and dL/dquery equal the
    reference's autograd (fixtures from the reference itself, make_golden_dcand.py), incl. a negative tied
    with the positive, a zero-norm candidate and explicit targets; a candidate-only gradient works too."""
    z, cfg_kw = load(golden_dir, case)
    cfg = xr.LossConfig(**cfg_kw)
    target = dev(z["target"]) if "target" in z else None
    for cls in xr.LOSS_CLASSES:
        name = cls.__name__
        q = dev(z["query"]).requires_grad_(True)
        cand = dev(z["cand"]).requires_grad_(True)
        loss = cls(cfg)(query_embed=q, candidate_embed=cand, target=target)
        assert float(loss) == pytest.approx(float(z[f"loss/{name}"]), rel=FP32_REL, abs=1e-5), (case, name)
        loss.backward()
        assert_close_grad(q.grad.cpu().numpy(), z[f"dq/{name}"], FP32_REL)
        got, want = cand.grad.cpu().numpy(), z[f"dcand/{name}"]
        assert got.shape == want.shape
        assert_close_grad(got.reshape(-1, got.shape[-1]), want.reshape(-1, want.shape[-1]), FP32_REL)
        cand2 = dev(z["cand"]).requires_grad_(True)
        (2.0 * cls(cfg)(query_embed=dev(z["query"]), candidate_embed=cand2, target=target)).backward()
        assert torch.allclose(cand2.grad, 2.0 * cand.grad, rtol=1e-6, atol=1e-9)
    with pytest.raises(NotImplementedError):    # handles stand for the frozen table: no gradient
        h = xr.PoolCandidates(dev(z["cand"])[:, 0], dev(z["cand"])[0])
        h.requires_grad = True
        xr.InfoNCELoss(xr.LossConfig())(query_embed=dev(z["query"]).requires_grad_(True), candidate_embed=h)


@pytest.mark.parametrize("case", GOLDEN_FP32)
def test_logits_statistics_match_reference(xr, golden_dir, case):
    z, cfg_kw = load(golden_dir, case)
    cfg = xr.LossConfig(**cfg_kw)
    target = dev(z["target"]) if "target" in z else None
    cand = dev(z["cand"]) if "cand" in z else xr.PoolCandidates(dev(z["pos"]), dev(z["neg"]))
    got = xr.LogitsStatistics(cfg)(query_embed=dev(z["query"]), candidate_embed=cand, target=target)
    want = json.loads(str(z["stats"]))
    assert set(got) == set(want)
    for k, v in want.items():
        if math.isnan(v):
            assert math.isnan(got[k]), k
        else:
            assert got[k] == pytest.approx(v, rel=2e-5, abs=2e-6), (case, k)


def test_losses_under_bf16_autocast(xr, golden_dir):
    """Lightning bf16-mixed (trainer.py:450): dot logits are bf16-rounded before masking."""
    z, cfg_kw = load(golden_dir, "pool_autocast_bf16")
    cfg = xr.LossConfig(**cfg_kw)
    cand = xr.PoolCandidates(dev(z["pos"]), dev(z["neg"]))
    for cls in xr.LOSS_CLASSES:
        name = cls.__name__
        q = dev(z["query"]).requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = cls(cfg)(query_embed=q, candidate_embed=cand)
        want = float(z[f"loss/{name}"])
        assert float(loss) == pytest.approx(want, rel=BF16_REL, abs=1e-4), name
        loss.backward()
        # gradients: compare with the oracle run on bf16-rounded logits (mask decided in bf16)
        _, dq, _, _ = orc.lean_loss(name, z["query"], z["pos"], z["neg"], orc.Config(**cfg_kw),
                                    with_grad=True, logits_dtype="bf16")
        ref = z[f"dq/{name}"]
        scale = max(float(np.abs(ref).max()), 1e-6)
        g = q.grad.float().cpu().numpy()
        # entries whose mask bit flipped under a 1-ulp bf16 difference are rare; judge in aggregate
        assert np.abs(g - dq).max() <= 0.05 * scale + 1e-6, name
        assert np.linalg.norm(g - dq) <= 1e-2 * np.linalg.norm(dq) + 1e-6, name


@pytest.mark.parametrize("name", orc.LOSS_NAMES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_sampled_candidates_match_oracle(xr, name, dtype):
    rng = np.random.default_rng(7)
    n, m, c, d = 500, 33, 65, 384
    table = (rng.standard_normal((n + 1, d)) / math.sqrt(d)).astype(np.float32)
    table[0] = 0
    if dtype == torch.bfloat16:
        table = orc.round_bf16(table)
    q = (rng.standard_normal((m, d)) / math.sqrt(d)).astype(np.float32)
    if dtype == torch.bfloat16:
        q = orc.round_bf16(q)
    idx = rng.integers(1, n + 1, size=(m, c))
    idx[3, 5] = idx[3, 0]      # duplicate of the positive among the negatives: tie -> masked
    idx[4, 7] = 0              # padding row as a candidate
    cfg_kw = {"margin": 0.3, "scale": 5.0}
    cand_dense = table[idx]
    logits_dtype = "bf16" if dtype == torch.bfloat16 else None
    want, dq = orc.embed_loss(name, q, cand_dense, orc.Config(**cfg_kw), with_grad=True,
                              logits_dtype=logits_dtype)
    qt = dev(q, dtype).requires_grad_(True)
    cand = xr.SampledCandidates(dev(table, dtype), dev(idx))
    loss = getattr(xr, name)(xr.LossConfig(**cfg_kw))(qt, cand)
    rel = FP32_REL if dtype == torch.float32 else BF16_REL
    assert float(loss) == pytest.approx(want, rel=rel * 2, abs=1e-4)
    loss.backward()
    g = qt.grad.float().cpu().numpy()
    tol = 4 * rel if dtype == torch.float32 else 2e-2
    assert np.abs(g - dq).max() <= tol * max(np.abs(dq).max(), 1e-6) + 1e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("cfg_kw", [{}, {"num_hard_negatives": 7}, {"mask_false_negatives": False, "scale": 4.0}])
def test_sampled_one_pass_equals_three_launches(xr, dtype, cfg_kw):
    """xr_sampled_step (logits + pipeline + dq in one launch, logits in shared memory) against
    xr_logits_sampled + xr_rowloss + xr_dq_sampled: same arithmetic, so the same bits."""
    _one_pass_vs_three(xr, dtype, cfg_kw, 5000, 700, 513)
    _one_pass_vs_three(xr, dtype, cfg_kw, 500, 33, 65)


def _one_pass_vs_three(xr, dtype, cfg_kw, n, m, c):
    g = torch.Generator(device="cuda").manual_seed(5)
    d = 384
    table = (torch.randn((n + 1, d), generator=g, device="cuda") / d ** 0.5).to(dtype)
    table[0] = 0
    q = (torch.randn((m, d), generator=g, device="cuda") / d ** 0.5).to(dtype)
    idx = torch.randint(0, n + 1, (m, c), generator=g, device="cuda")
    idx[3, 5] = idx[3, 0]
    for name in orc.LOSS_NAMES:
        res = []
        for one_pass in (True, False):
            xr.losses._SAMPLED_ONE_PASS = one_pass
            try:
                qt = q.clone().requires_grad_(True)
                loss = getattr(xr, name)(xr.LossConfig(**cfg_kw))(qt, xr.SampledCandidates(table, idx))
                loss.backward()
                res.append((loss.detach().clone(), qt.grad.clone()))
            finally:
                xr.losses._SAMPLED_ONE_PASS = True
        assert torch.equal(res[0][0], res[1][0]), name
        assert torch.equal(res[0][1], res[1][1]), name
    stats1 = xr.LogitsStatistics(xr.LossConfig(**cfg_kw))(q, xr.SampledCandidates(table, idx))
    xr.losses._SAMPLED_ONE_PASS = False
    try:
        stats3 = xr.LogitsStatistics(xr.LossConfig(**cfg_kw))(q, xr.SampledCandidates(table, idx))
    finally:
        xr.losses._SAMPLED_ONE_PASS = True
    assert stats1 == stats3


def test_pool_large_fp32_matches_oracle(xr):
    """BASELINE config-1 shape class (fp32, shared pool), rows > one GEMM tile, ragged sizes."""
    rng = np.random.default_rng(11)
    m, cn, d = 301, 777, 384
    q = (rng.standard_normal((m, d)) / math.sqrt(d)).astype(np.float32)
    pos = (rng.standard_normal((m, d)) / math.sqrt(d)).astype(np.float32)
    neg = (rng.standard_normal((cn, d)) / math.sqrt(d)).astype(np.float32)
    for name in orc.LOSS_NAMES:
        want, dq, _, _ = orc.lean_loss(name, q, pos, neg, orc.Config(), with_grad=True)
        qt = dev(q).requires_grad_(True)
        loss = getattr(xr, name)(xr.LossConfig())(qt, xr.PoolCandidates(dev(pos), dev(neg)))
        assert float(loss) == pytest.approx(want, rel=FP32_REL, abs=1e-5), name
        loss.backward()
        assert_close_grad(qt.grad.cpu().numpy(), dq, FP32_REL)


def test_loss_error_behaviour(xr):
    q = torch.randn(4, 8).cuda()
    cand = torch.randn(4, 5, 8).cuda()
    with pytest.raises(AssertionError):
        xr.InfoNCELoss(xr.LossConfig())(q[0], cand)                       # losses.py:166
    with pytest.raises(AssertionError):
        xr.InfoNCELoss(xr.LossConfig())(q, cand[:, 0])                    # losses.py:169
    with pytest.raises(AssertionError):
        xr.InfoNCELoss(xr.LossConfig())(q[:3], cand)                      # losses.py:172
    with pytest.raises(AssertionError):
        xr.InfoNCELoss(xr.LossConfig())(q, cand, target=torch.zeros(4).long().cuda())  # :236
    with pytest.raises(AssertionError):
        xr.InfoNCELoss(xr.LossConfig(target_position=None))(q, cand)      # losses.py:233
    with pytest.raises(IndexError):
        xr.InfoNCELoss(xr.LossConfig(target_position=None))(q, cand, target=torch.full((4,), 9).cuda())
    cfg = xr.LossConfig()
    object.__setattr__(cfg, "target_position", "bogus")
    with pytest.raises(ValueError):
        xr.InfoNCELoss(cfg)(q, cand)                                       # losses.py:251-253
    assert len(list(xr.InfoNCELoss(xr.LossConfig()).parameters())) == 0
    assert len(xr.InfoNCELoss(xr.LossConfig()).state_dict()) == 0          # checkpoint keys unchanged
    with pytest.raises(xr._native.NativeError):
        xr.InfoNCELoss(xr.LossConfig())(q.cpu(), cand.cpu())               # no CPU fallback


# ------------------------------------------------------------------------------------------
# family 3: top-k, exact search, metrics
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("u,n,k", [(1, 5, 2), (3, 1000, 100), (7, 70001, 100), (1, 300000, 100),
                                   (130, 4099, 20), (2, 50, 100), (5, 2048, 1024)])
def test_topk_bit_exact_vs_stable_sort(xr, u, n, k):
    rng = np.random.default_rng(u * 1000 + k)
    s = rng.standard_normal((u, n)).astype(np.float32)
    s = np.round(s * 8) / 8            # heavy ties: the (score desc, index asc) rule decides
    if n > 10:
        s[0, 3] = np.inf
        s[0, 7] = -np.inf
        s[-1, 5] = -0.0
        s[-1, 6] = 0.0
    got_s, got_i = xr.ops.topk(dev(s), k)
    want_s, want_i = orc.topk_rows(s, k)
    kk = min(k, n)
    assert np.array_equal(got_i.cpu().numpy()[:, :kk], want_i)
    assert np.array_equal(got_s.cpu().numpy()[:, :kk], want_s)
    if k > n:
        assert (got_i.cpu().numpy()[:, n:] == -1).all()
        assert np.isneginf(got_s.cpu().numpy()[:, n:]).all()


def test_topk_sorted_and_adversarial_inputs(xr):
    n, k = 100000, 100
    asc = np.arange(n, dtype=np.float32)[None]          # every element beats the threshold
    _, i = xr.ops.topk(dev(asc), k)
    assert np.array_equal(i.cpu().numpy()[0], np.arange(n - 1, n - 1 - k, -1))
    const = np.zeros((2, n), np.float32)                # all tied: lowest ids win
    _, i = xr.ops.topk(dev(const), k)
    assert np.array_equal(i.cpu().numpy(), np.tile(np.arange(k), (2, 1)))


def test_topk_merge_equals_global_topk(xr):
    rng = np.random.default_rng(3)
    u, n, k, g = 9, 8000, 100, 4
    s = np.round(rng.standard_normal((u, n)).astype(np.float32) * 4) / 4
    per = n // g
    ps, pi = [], []
    for r in range(g):
        a, b = xr.ops.topk(dev(s[:, r * per:(r + 1) * per].copy()), k, col_offset=r * per)
        ps.append(a)
        pi.append(b)
    ms, mi = xr.ops.topk_merge(torch.cat(ps, 1), torch.cat(pi, 1), k)
    want_s, want_i = orc.topk_rows(s, k)
    assert np.array_equal(mi.cpu().numpy(), want_i)
    assert np.array_equal(ms.cpu().numpy(), want_s)


@pytest.mark.parametrize("metric", ["cosine", "dot"])
def test_exact_index_matches_oracle(xr, metric):
    """Small-integer embeddings make every dot product exact in bf16/fp32, so indices must be
    IDENTICAL to the stable-sort oracle, ties included."""
    rng = np.random.default_rng(5)
    n, d, u, k = 5000, 384, 6, 100
    cat = rng.integers(-2, 3, size=(n, d)).astype(np.float32)
    qs = rng.integers(-2, 3, size=(u, d)).astype(np.float32)
    excl = [list(rng.integers(0, n, size=rng.integers(20, 200))) for _ in range(u)]
    idx = xr.index.ExactIndex(xr.index.ExactIndexConfig(index_metric=metric, dtype="fp32",
                                                        max_score_bytes=u * 4 * 2048))
    idx.set_catalog(torch.from_numpy(cat).cuda())
    s, i = idx.search_batch(dev(qs), excl, k)
    if metric == "dot":
        want_s, want_i = orc.exact_search(qs, cat, k, excl, metric="dot")
        assert np.array_equal(i.cpu().numpy(), want_i)
        assert np.array_equal(s.cpu().numpy(), want_s)
    else:
        want_s, want_i = orc.exact_search(qs, cat, k, excl, metric="cosine", dtype=np.float64)
        got_i, got_s = i.cpu().numpy(), s.cpu().numpy()
        # normalisation is not exact arithmetic: require the same scores, and identical ids
        # wherever the oracle's neighbouring scores are separated by more than fp32 noise
        assert np.allclose(got_s, want_s, atol=2e-6)
        gap = np.abs(np.diff(want_s, axis=1)).min(axis=1)
        for r in range(u):
            if gap[r] > 1e-5:
                assert np.array_equal(got_i[r], want_i[r])
            assert not set(got_i[r].tolist()) & set(int(x) for x in excl[r])


def test_index_search_reference_surface(xr):
    import datasets

    rng = np.random.default_rng(9)
    n, d = 300, 384
    emb = rng.standard_normal((n, d)).astype(np.float32)
    ds = datasets.Dataset.from_dict({"item_id": [f"i{j}" for j in range(n)],
                                     "item_text": [f"text {j}" for j in range(n)],
                                     "embedding": emb.tolist()})
    idx = xr.index.ExactIndex(xr.index.ExactIndexConfig(dtype="fp32")).index_data(ds)
    q = emb[17] + 0.01 * rng.standard_normal(d).astype(np.float32)
    res = idx.search(q, exclude_item_ids=["i17", "i3", "nope"], top_k=20)
    assert len(res) == 20 and {"item_id", "score", "item_text"} <= set(res.column_names)
    assert "i17" not in res["item_id"] and "i3" not in res["item_id"]
    want_s, want_i = orc.exact_search(q, emb, 20, [[17, 3]], dtype=np.float64)
    assert res["item_id"] == [f"i{j}" for j in want_i[0]]
    assert np.allclose(res["score"], want_s[0], atol=1e-5)
    assert idx.get_id("i5")["item_text"] == "text 5" and idx.get_id(None) == {} and idx.get_id("zz") == {}
    res = idx.search(q, None, top_k=400)          # more than the catalog holds
    assert len(res) == n


def test_retrieval_metrics_match_oracle(xr):
    rng = np.random.default_rng(2)
    for top_k in (1, 5, 20, 100):
        recs, tgts = [], []
        for _ in range(40):
            k = int(rng.integers(0, top_k + 3))
            rec = [int(x) for x in rng.choice(60, size=min(k, 60), replace=False)]
            tgt = [int(x) for x in rng.choice(60, size=int(rng.integers(0, 11)), replace=False)]
            recs.append(rec)
            tgts.append(tgt)
        width = max(top_k, max(len(r) for r in recs))
        mat = torch.full((len(recs), width), -1, dtype=torch.int64)
        for r, rec in enumerate(recs):
            mat[r, :len(rec)] = torch.tensor(rec, dtype=torch.int64) if rec else mat[r, :0]
        out, valid = xr.metrics.retrieval_metrics_batch(mat.cuda(), tgts, top_k)
        out, valid = out.cpu().numpy(), valid.cpu().numpy()
        for r, (rec, tgt) in enumerate(zip(recs, tgts)):
            want = orc.retrieval_metrics([str(x) for x in rec], {str(x) for x in tgt}, top_k)
            assert bool(valid[r]) == (len(want) > 0)
            for c, name in enumerate(orc.METRIC_NAMES):
                if want:
                    assert out[r, c] == pytest.approx(want[name], rel=1e-6, abs=1e-7), (name, rec, tgt)
    m = xr.metrics.compute_retrieval_metrics(["i10", "i3", "i7"], {"i3"}, top_k=3)
    assert list(m) == orc.METRIC_NAMES and float(m["retrieval_auroc"]) == pytest.approx(0.5)
    assert xr.metrics.compute_retrieval_metrics(["a"], set(), 3) == {}


def test_config1_shape_full_size_matches_oracle(xr):
    """BASELINE configs[0] at full size: ML-1M-shaped (3,706 items), B=128, L=50, fp32, InfoNCE with
    the in-batch shared pool — compute_embeds + loss + backward against the numpy oracle
    (the lean oracle at B=128; the verbatim (M, 1+M, D) tensor would need 2 x 63 GB)."""
    b = orc.synth_batch(3706, 128, 50, dim=384, seed=0)
    emb = xr.models.ItemEmbeddings(torch.from_numpy(b["table"]), add_padding_row=False).cuda()
    tok = torch.from_numpy(b["token_embeddings"]).cuda().requires_grad_(True)
    hist, pos, neg = (torch.from_numpy(b[k]).cuda() for k in
                      ("history_item_idx", "pos_item_idx", "neg_item_idx"))
    out = xr.models.compute_embeds(emb, tok, hist, pos, neg)
    want = orc.compute_embeds(b["table"], b["token_embeddings"], b["history_item_idx"],
                              b["pos_item_idx"], b["neg_item_idx"], dense=False)
    assert np.array_equal(out["query_embed"].detach().cpu().numpy(), want["query_embed"])
    assert np.array_equal(out["candidate_embed"].neg.cpu().numpy(), want["neg_embed"])
    loss = xr.InfoNCELoss(xr.LossConfig())(out["query_embed"], out["candidate_embed"])
    loss.backward()
    ref, dq, _, _ = orc.lean_loss("InfoNCELoss", want["query_embed"], want["pos_embed"],
                                  want["neg_embed"], orc.Config(), with_grad=True)
    assert float(loss) == pytest.approx(ref, rel=FP32_REL)
    mask = (b["history_item_idx"].reshape(-1) != 0) & (b["pos_item_idx"].reshape(-1) != 0)
    g = tok.grad.reshape(-1, 384).cpu().numpy()
    assert not g[~mask].any()
    assert_close_grad(g[mask], dq, FP32_REL)


def test_evaluate_batch_equals_per_user_loop(xr):
    """the batched validation loop gives, user by user, what the reference's
    validation_step computes (search with the history excluded, then the 7 metrics), and the
    epoch means over the users that have targets."""
    rng = np.random.default_rng(5)
    n, u, k = 5000, 37, 20
    cat = rng.standard_normal((n, 384)).astype(np.float32)
    qs = rng.standard_normal((u, 384)).astype(np.float32)
    hist = [list(map(int, rng.integers(0, n, size=int(rng.integers(0, 40))))) for _ in range(u)]
    tgts = [list(map(int, rng.integers(0, n, size=int(rng.integers(0, 6))))) for _ in range(u)]
    tgts[3] = []                                   # a user without targets: no metrics logged
    ws, wi = orc.exact_search(qs, cat, k, hist, metric="cosine")
    tgts[5] = [int(wi[5, 0]), int(wi[5, 7])]       # guaranteed hits
    idx = xr.index.ExactIndex(xr.index.ExactIndexConfig(index_metric="cosine", dtype="fp32")).set_catalog(
        torch.from_numpy(cat).cuda())
    means, per_user, valid, rec = xr.evaluate.evaluate_batch(idx, torch.from_numpy(qs).cuda(), hist, tgts, k)
    assert np.array_equal(rec.cpu().numpy(), wi)
    names = xr.metrics.METRIC_NAMES
    acc = np.zeros(7)
    cnt = 0
    for r in range(u):
        want = orc.retrieval_metrics([str(x) for x in wi[r]], {str(x) for x in tgts[r]}, k)
        assert bool(valid[r]) == bool(want)
        if want:
            got = per_user[r].cpu().numpy()
            for j, name in enumerate(names):
                assert got[j] == pytest.approx(want[name], rel=1e-5, abs=1e-6), (r, name)
            acc += [want[name] for name in names]
            cnt += 1
    for j, name in enumerate(names):
        assert float(means[f"val/{name}"]) == pytest.approx(acc[j] / cnt, rel=1e-5, abs=1e-6)


def test_search_batch_query_blocks_equal_one_shot(xr):
    """Large query sets run in blocks that bound the group-maxima buffer: same result."""
    rng = np.random.default_rng(3)
    n, u, k = 20000, 700, 50
    cat = torch.from_numpy(rng.standard_normal((n, 384)).astype(np.float32)).cuda()
    q = torch.from_numpy(rng.standard_normal((u, 384)).astype(np.float32)).cuda()
    excl = [list(map(int, rng.integers(0, n, size=int(rng.integers(0, 30))))) for _ in range(u)]
    one = xr.index.ExactIndex(xr.index.ExactIndexConfig()).set_catalog(cat)
    blk = xr.index.ExactIndex(xr.index.ExactIndexConfig(max_groupmax_bytes=1)).set_catalog(cat)   # 256-query blocks
    for ex in (excl, xr.ops._csr(excl, cat.device), None):
        s1, i1 = one.search_batch(q, ex, k)
        s2, i2 = blk.search_batch(q, ex, k)
        assert torch.equal(i1, i2) and torch.equal(s1, s2)


def test_index_table_round_trip(xr, tmp_path):
    """the items table in the reference's layout ({item_id, item_text, embedding:
    fixed_size_list<float32>[D]}, Parquet) written by save_table and reopened by load_table gives the
    same search results; a table produced WITHOUT this package (plain pyarrow, as the reference's data
    pipeline writes it) opens too."""
    import pyarrow as pa
    import pyarrow.parquet as pq

    rng = np.random.default_rng(8)
    n, d, k = 3000, 384, 20
    emb = rng.standard_normal((n, d)).astype(np.float32)
    data = {"item_id": [f"m{i}" for i in range(n)], "item_text": [f"title {i}" for i in range(n)],
            "embedding": emb}
    idx = xr.index.ExactIndex(xr.index.ExactIndexConfig(dtype="fp32")).index_data(data)
    q = torch.from_numpy(rng.standard_normal((5, d)).astype(np.float32)).cuda()
    s0, i0 = idx.search_batch(q, [[1, 2], [], [7], [], []], k)
    p = tmp_path / "items.parquet"
    idx.save_table(str(p), embeddings=torch.from_numpy(emb))
    back = xr.index.ExactIndex.load_table(str(p))
    assert back.config == idx.config and back.ids == idx.ids and back.columns == idx.columns
    s1, i1 = back.search_batch(q, [[1, 2], [], [7], [], []], k)
    assert torch.equal(i0, i1) and torch.equal(s0, s1)
    one = back.search(emb[11], exclude_item_ids=["m11"], top_k=5)
    assert "m11" not in one["item_id"] and one["item_text"][0].startswith("title ")
    # a foreign table with the reference's columns
    t = pa.table({"item_id": pa.array(data["item_id"]), "item_text": pa.array(data["item_text"]),
                  "embedding": pa.FixedSizeListArray.from_arrays(pa.array(emb.reshape(-1)), d)})
    p2 = tmp_path / "foreign.parquet"
    pq.write_table(t, str(p2))
    fx = xr.index.ExactIndex.load_table(str(p2), xr.index.ExactIndexConfig(dtype="fp32"))
    s2, i2 = fx.search_batch(q, [[1, 2], [], [7], [], []], k)
    assert torch.equal(i0, i2) and torch.equal(s0, s2)


@pytest.mark.parametrize("case", ["default", "scale_margin_nomask"])
def test_compute_losses_matches_the_reference_trainer(xr, golden_dir, case):
    """The whole training-step log dict against the reference's OWN
    RecommenderLightningModule.compute_losses (trainer.py:213-264) run on its own compute_embeds and
    loss classes (tests/golden/make_golden_compute_losses.py): same 30 keys in the same order, fp32
    values within 1e-5, and the gradient of the train loss at the encoder output."""
    z = np.load(golden_dir / f"compute_losses_{case}.npz")
    cfg = xr.LossConfig(**json.loads(str(z["cfg"])))
    want = json.loads(str(z["logged"]))
    emb = xr.models.ItemEmbeddings(torch.from_numpy(z["table"]), add_padding_row=False).cuda()
    tok = torch.from_numpy(z["tokens"]).cuda().requires_grad_(True)
    args = [torch.from_numpy(z[k]).cuda() for k in ("history_item_idx", "pos_item_idx", "neg_item_idx")]
    for dense in (False, True):
        tok.grad = None
        embeds = xr.models.compute_embeds(emb, tok, *args, dense=dense)
        got = xr.losses.compute_losses(cfg, embeds)
        assert list(got.keys()) == json.loads(str(z["keys"]))
        for k, v in want.items():
            g = float(got[k])
            if v != v:
                assert g != g, k
            else:
                assert g == pytest.approx(v, rel=FP32_REL, abs=1e-6), (k, dense)
        got["loss/InfoNCELoss"].backward()
        assert_close_grad(tok.grad.cpu().numpy(), z["dtokens"], FP32_REL)


@pytest.mark.parametrize("u", [1, 7, 200])
def test_compiled_search_plan_equals_search_batch(xr, u):
    """ExactIndex.compile_search: the whole search as one CUDA-graph replay, same result as
    search_batch for every replay (different queries / exclusion lists through the static buffers)."""
    if torch.cuda.get_device_capability()[0] != 10:
        pytest.skip("needs an sm_100 device")
    rng = np.random.default_rng(u)
    n, k = 50_000, 20
    cat = torch.from_numpy(rng.standard_normal((n, 384)).astype(np.float32)).cuda()
    idx = xr.index.ExactIndex(xr.index.ExactIndexConfig()).set_catalog(cat)
    plan = idx.compile_search(u, k, max_exclusions=40)
    plain = idx.compile_search(u, k)
    for rep in range(3):
        q = torch.from_numpy(rng.standard_normal((u, 384)).astype(np.float32)).cuda()
        excl = [list(map(int, rng.integers(0, n, size=int(rng.integers(0, 41))))) for _ in range(u)]
        csr = xr.ops._csr(excl, cat.device)
        want_s, want_i = idx.search_batch(q, csr, k)
        got_s, got_i = plan(q, csr)
        assert torch.equal(got_i, want_i) and torch.equal(got_s, want_s)
        want_s, want_i = idx.search_batch(q, None, k)
        got_s, got_i = plan(q, None)
        assert torch.equal(got_i, want_i) and torch.equal(got_s, want_s)
        got_s, got_i = plain(q)
        assert torch.equal(got_i, want_i) and torch.equal(got_s, want_s)


def test_torch_library_ops_match_the_wrappers(xr):
    """torch.ops.xfmr_b200.* (SURVEY 8b): same results as the ctypes wrappers; ::pool_loss carries the
    autograd edge (backward = the gradient the fused kernel produced, scaled by grad_output)."""
    if torch.cuda.get_device_capability()[0] != 10:
        pytest.skip("needs an sm_100 device")
    from xfmr_rec_b200 import _native as N, ops

    g = torch.Generator(device="cuda").manual_seed(5)
    table = torch.randn((500, 384), generator=g, device="cuda")
    idx = torch.randint(0, 500, (7, 9), generator=g, device="cuda")
    assert torch.equal(torch.ops.xfmr_b200.gather_rows(table, idx), table[idx])
    q = (torch.randn((300, 384), generator=g, device="cuda") / 20).bfloat16()
    pos = (torch.randn((300, 384), generator=g, device="cuda") / 20).bfloat16()
    neg = (torch.randn((777, 384), generator=g, device="cuda") / 20).bfloat16()
    kind = N.LOSS_KIND["InfoNCELoss"]
    loss, dq = torch.ops.xfmr_b200.score_loss_fwd_bwd(q, pos, neg, kind, True, 1.0, 0.5, True)
    want = xr.InfoNCELoss(xr.LossConfig())(q.clone().requires_grad_(True), xr.PoolCandidates(pos, neg))
    assert torch.equal(loss, want.detach())
    qg = q.clone().requires_grad_(True)
    l2, _ = torch.ops.xfmr_b200.pool_loss(qg, pos, neg, kind, True, 1.0, 0.5, True)
    (3.0 * l2).backward()
    assert torch.equal(qg.grad, (dq * 3.0).to(torch.bfloat16))
    sc = torch.randn((5, 3000), generator=g, device="cuda")
    s1, i1 = torch.ops.xfmr_b200.topk(sc, 10)
    s2, i2 = ops.topk(sc, 10)
    assert torch.equal(s1, s2) and torch.equal(i1, i2)
    cat, _ = ops.normalize_rows(torch.randn((20000, 384), generator=g, device="cuda"), 1e-12, torch.bfloat16)
    qq, _ = ops.normalize_rows(torch.randn((9, 384), generator=g, device="cuda"), 1e-12, torch.bfloat16)
    s3, i3 = torch.ops.xfmr_b200.score_topk(qq, cat, 20)
    idx2 = xr.index.ExactIndex(xr.index.ExactIndexConfig(index_metric="dot")).set_catalog(cat)
    s4, i4 = idx2.search_batch(qq, None, 20)
    assert torch.equal(i3, i4) and torch.equal(s3, s4)
    offs = torch.tensor([0, 2, 3], device="cuda")
    ids = torch.tensor([int(i3[0, 1]), 5, int(i3[1, 0])], device="cuda")
    m1, v1 = torch.ops.xfmr_b200.retrieval_metrics(i3[:2].contiguous(), offs, ids, 20)
    m2, v2 = ops.retrieval_metrics(i3[:2].contiguous(), (offs, ids), 20)
    assert torch.equal(m1, m2) and torch.equal(v1, v2)


def test_item_index_service_wire_format(xr):
    """service.py:137-180 over ExactIndex: Query -> list[ItemCandidate] equals index.search, batched
    search_many equals the per-query calls, get_id / get_ids return ItemQuery rows (embedding included,
    the ORIGINAL rows with store_embeddings) and unknown ids raise NotFound."""
    from xfmr_rec_b200 import service as S

    rng = np.random.default_rng(2)
    n = 3000
    emb = rng.standard_normal((n, 384)).astype(np.float32)
    data = {"item_id": [f"i{j}" for j in range(n)], "item_text": [f"text {j}" for j in range(n)],
            "embedding": torch.from_numpy(emb)}
    idx = xr.index.ExactIndex(xr.index.ExactIndexConfig(store_embeddings=True)).index_data(data)
    svc = S.ItemIndexService(idx)
    qs = [S.Query(embedding=rng.standard_normal(384).astype(np.float32), exclude_item_ids=[f"i{j}" for j in range(r * 3)],
                  top_k=5 + r) for r in range(4)]
    many = svc.search_many(qs)
    for q, got in zip(qs, many):
        one = svc.search(q)
        ref = idx.search(q.embedding, q.exclude_item_ids, q.top_k)
        assert [c.item_id for c in got] == [c.item_id for c in one] == list(ref["item_id"])
        assert len(got) == q.top_k and all(isinstance(c, S.ItemCandidate) for c in got)
        assert [c.item_text for c in got] == list(ref["item_text"])
        assert not set(c.item_id for c in got) & set(q.exclude_item_ids or [])
        np.testing.assert_allclose([c.score for c in got], ref["score"], rtol=1e-6)
    item = svc.get_id("i7")
    assert item.item_id == "i7" and item.item_text == "text 7"
    np.testing.assert_array_equal(item.embedding, emb[7])
    assert set(svc.get_ids(["i1", "i2", "nope"])) == {"i1", "i2"}
    with pytest.raises(S.NotFound):
        svc.get_id("nope")
    wire = S.Query.model_validate_json(qs[1].model_dump_json())       # JSON round trip of a request
    assert [c.item_id for c in svc.search(wire)] == [c.item_id for c in many[1]]


def test_recommend_service_request_path(xr):
    """service.py:96-134 + 202-262 without BentoML: item ids -> stored embeddings -> Model.embed (the B200-native
    sequence encoder, eval mode) -> search with the query's own items excluded.  The batched entry point equals the
    per-query calls; the embedding equals the encoder's pooled output for the same item rows; unknown ids are
    dropped, an empty query returns no candidates."""
    from xfmr_rec_b200 import service as S
    from xfmr_rec_b200.encoder import EncoderConfig, SeqEncoder

    rng = np.random.default_rng(4)
    n = 2000
    emb = (rng.standard_normal((n, 384)) / 384 ** 0.5).astype(np.float32)
    data = {"item_id": [f"i{j}" for j in range(n)], "item_text": [f"text {j}" for j in range(n)],
            "embedding": torch.from_numpy(emb)}
    idx = xr.index.ExactIndex(xr.index.ExactIndexConfig(store_embeddings=True)).index_data(data)
    torch.manual_seed(0)
    enc = SeqEncoder(EncoderConfig(num_hidden_layers=2, intermediate_size=96, max_seq_length=8)).cuda()
    svc = S.RecommendService(S.ModelService(enc), S.ItemIndexService(idx))
    assert not enc.training
    hists = [[f"i{j}" for j in rng.integers(0, n, size=k)] for k in (3, 12, 1)]
    hists[1][2] = "unknown-item"
    qs = [S.Query(item_ids=list(h), top_k=7, exclude_item_ids=["i5"]) for h in hists] + [S.Query(top_k=3)]
    many = svc.recommend_with_queries([q.model_copy(deep=True) for q in qs])
    for q, got, h in zip(qs, many, hists + [[]]):
        one = svc.recommend_with_query(q.model_copy(deep=True))
        assert [c.item_id for c in one] == [c.item_id for c in got]
        if not h:
            assert got == []
            continue
        assert len(got) == 7 and not {c.item_id for c in got} & ({"i5"} | set(h))
    # the embedding of query 1: the encoder on the last 8 KNOWN items of its history (table = the item embeddings)
    known = [int(x[1:]) for x in hists[1] if x != "unknown-item"][-8:]
    table = torch.cat([torch.zeros(1, 384), torch.from_numpy(emb)]).cuda()
    want = enc(torch.tensor([[r + 1 for r in known]], device="cuda"), table)["sentence_embedding"][0]
    q1 = svc.embed_query(svc.process_query(qs[1].model_copy(deep=True)))
    np.testing.assert_allclose(q1.embedding, want.detach().cpu().numpy(), rtol=1e-5, atol=1e-6)
    first = svc.recommend_with_item_id("i3", top_k=4)
    assert len(first) == 4 and "i3" not in {c.item_id for c in first}
    users = S.UserIndexService([{"user_id": "u1", "user_text": "{}", "history": {"item_id": ["i1", "i2"], "item_text": ["a", "b"]},
                                 "target": {"item_id": ["i9"], "item_text": ["c"]}}])
    recs = svc.recommend_with_user_id(users, "u1", top_k=6)
    want = svc.recommend_with_query(S.Query(item_ids=["i1", "i2", "i9"], top_k=6))
    assert [c.item_id for c in recs] == [c.item_id for c in want] and not {"i1", "i2", "i9"} & {c.item_id for c in recs}
    with pytest.raises(S.NotFound):
        users.get_id("nobody")
