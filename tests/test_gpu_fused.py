"""GPU parity tests for the tcgen05/TMEM fused contraction+loss kernel (xr_fused_pool_loss)
against the oracle on bf16-rounded inputs and against the library's own fp32-accumulate
materialised path."""

import math

import numpy as np
import pytest
import torch

from oracle import xfmr_oracle as orc

pytestmark = pytest.mark.gpu

FUSED = ["InfoNCELoss", "NCELoss", "PairwiseHingeLoss", "PairwiseLogisticLoss", "ContrastiveLoss",
         "AlignmentContrastiveLoss"]


@pytest.fixture(scope="module")
def xr():
    import xfmr_rec_b200 as pkg

    if not torch.cuda.is_available() or torch.cuda.get_device_capability()[0] != 10:
        pytest.skip("needs an sm_100 device")
    assert pkg._native.lib().xr_fused_available() & 1, "tcgen05 kernels missing from the build"
    return pkg


def make_inputs(m, cn, d=384, seed=0, dup=True, scale_q=1.0):
    rng = np.random.default_rng(seed)
    s = 1.0 / math.sqrt(d)
    q = orc.round_bf16((rng.standard_normal((m, d)) * s * scale_q).astype(np.float32))
    pos = orc.round_bf16((rng.standard_normal((m, d)) * s).astype(np.float32))
    neg = orc.round_bf16((rng.standard_normal((cn, d)) * s).astype(np.float32))
    if dup and cn > 3 and m > 2:
        neg[1] = pos[0]          # the row's own positive inside the pool: exact tie -> masked
        neg[cn - 1] = pos[m - 1]
        neg[2] = 0.0             # a padding row in the pool
    return q, pos, neg


def bf(x):
    return torch.from_numpy(x).cuda().bfloat16()


def run_fused(xr, name, q, pos, neg, cfg_kw, logits_bf16, want_grad=True):
    from xfmr_rec_b200 import _native as N, ops

    class C:
        pass

    c = C()
    for k, v in {**dict(mask_false_negatives=True, num_hard_negatives=0, scale=1.0, margin=0.5), **cfg_kw}.items():
        setattr(c, k, v)
    cosine = name in orc.COSINE_LOSSES
    cfg = ops.make_cfg(c, logits_bf16=logits_bf16 and not cosine)
    qt, pt, nt = bf(q), bf(pos), bf(neg)
    q_inv = None
    if cosine:
        qt, q_inv = ops.normalize_rows(torch.from_numpy(q).cuda(), 1e-8, torch.bfloat16)
        pt, _ = ops.normalize_rows(torch.from_numpy(pos).cuda(), 1e-8, torch.bfloat16)
        nt, _ = ops.normalize_rows(torch.from_numpy(neg).cuda(), 1e-8, torch.bfloat16)
    loss, dq, _ = ops.fused_pool_loss(qt, pt, nt, N.LOSS_KIND[name], cfg, q_inv=q_inv,
                                      want_grad=want_grad)
    torch.cuda.synchronize()
    extras = (qt.float().cpu().numpy(), pt.float().cpu().numpy(), nt.float().cpu().numpy(),
              None if q_inv is None else q_inv.cpu().numpy())
    return float(loss[0]), (None if dq is None else dq.cpu().numpy()), extras


def oracle_on_bf16(name, qh, ph, nh, q_inv, cfg_kw, logits_bf16):
    """Oracle on exactly the bf16 operands the kernel saw.  For the cosine kinds those are the
    already-normalised rows: logits are plain dots of them, and d/dq chains through q_inv."""
    cfg = orc.Config(**cfg_kw)
    cosine = name in orc.COSINE_LOSSES
    logits = orc.lean_logits(qh, ph, nh)
    if logits_bf16 and not cosine:
        logits = orc.round_bf16(logits.astype(np.float32)).astype(np.float64)
    else:
        logits = logits.astype(np.float32).astype(np.float64)
    tgt = np.zeros(qh.shape[0], np.int64)
    mask = orc.mask_false_negatives(logits, tgt, cfg)
    loss, g = orc.loss_from_logits(name, logits, tgt, mask, cfg, with_grad=True)
    ghat = g[:, :1] * ph + g[:, 1:] @ nh
    if cosine:
        ghat = q_inv[:, None] * (ghat - (ghat * qh).sum(-1, keepdims=True) * qh)
    return loss, ghat


@pytest.mark.parametrize("name", FUSED)
@pytest.mark.parametrize("m,cn", [(128, 64), (1, 1), (130, 65), (301, 777), (700, 3000)])
def test_fused_matches_oracle(xr, name, m, cn):
    q, pos, neg = make_inputs(m, cn, seed=m + cn)
    for cfg_kw, lbf in [({}, True), ({"scale": 8.0, "margin": 0.2}, True), ({"mask_false_negatives": False}, False)]:
        loss, dq, (qh, ph, nh, q_inv) = run_fused(xr, name, q, pos, neg, cfg_kw, lbf)
        want, want_dq = oracle_on_bf16(name, qh, ph, nh, q_inv, cfg_kw, lbf)
        assert loss == pytest.approx(want, rel=2e-3, abs=2e-3), (name, cfg_kw, loss, want)
        scale = max(np.abs(want_dq).max(), 1e-6)
        # north_star: 2e-3 in bf16.  The kernel rounds the softmax / sigmoid weights to bf16 for the
        # gradient MMA (2^-9 per weight, averaging out over the candidates); measured 1.4e-4 .. 4.2e-4
        # norm-wise, <= 1.8e-3 of the largest element (profiles/grad_error_r02.json) -- the reference's
        # own bf16-autocast gradient is 2.0e-3 .. 2.6e-3 / 3.6e-3 .. 4.1e-3 from the same oracle
        assert np.linalg.norm(dq - want_dq) <= 2e-3 * np.linalg.norm(want_dq) + 1e-6, (name, cfg_kw)
        assert np.abs(dq - want_dq).max() <= 6e-3 * scale + 1e-6, (name, cfg_kw)


def test_fused_forward_only_equals_forward_backward(xr):
    q, pos, neg = make_inputs(260, 900, seed=3)
    for name in FUSED:
        a, _, _ = run_fused(xr, name, q, pos, neg, {}, True, want_grad=True)
        b, dq, _ = run_fused(xr, name, q, pos, neg, {}, True, want_grad=False)
        assert dq is None and a == b, name


def test_fused_is_deterministic(xr):
    q, pos, neg = make_inputs(513, 2100, seed=4)
    a, da, _ = run_fused(xr, "InfoNCELoss", q, pos, neg, {}, True)
    b, db, _ = run_fused(xr, "InfoNCELoss", q, pos, neg, {}, True)
    assert a == b and np.array_equal(da, db)


def test_fused_duplicate_positive_is_masked(xr):
    """In-batch negatives routinely contain the row's own positive item; the reference masks it
    because one bmm gives both logits identical bits (losses.py:289-292)."""
    m, cn = 256, 256
    q, pos, neg = make_inputs(m, cn, seed=9, dup=False)
    neg[:] = pos            # every row's positive sits in the pool (the in-batch setting)
    loss, dq, (qh, ph, nh, _) = run_fused(xr, "InfoNCELoss", q, pos, neg, {}, False)
    want, _ = oracle_on_bf16("InfoNCELoss", qh, ph, nh, None, {}, False)
    assert loss == pytest.approx(want, rel=1e-4)


def test_loss_modules_take_the_fused_path(xr):
    """Through the public loss modules: bf16 inputs + PoolCandidates -> tcgen05 kernel."""
    q, pos, neg = make_inputs(300, 1000, seed=5)
    for name in orc.LOSS_NAMES:
        qt = bf(q).requires_grad_(True)
        cand = xr.PoolCandidates(bf(pos), bf(neg))
        loss = getattr(xr, name)(xr.LossConfig())(qt, cand)
        loss.backward()
        lb = None if name in orc.COSINE_LOSSES else "bf16"
        want, want_dq, _, _ = orc.lean_loss(name, q, pos, neg, orc.Config(), with_grad=True, logits_dtype=lb)
        assert float(loss) == pytest.approx(want, rel=4e-3, abs=4e-3), name
        g = qt.grad.float().cpu().numpy()
        assert np.linalg.norm(g - want_dq) <= 2e-2 * np.linalg.norm(want_dq) + 1e-6, name


def test_fused_under_autocast_matches_golden(xr, golden_dir):
    """fp32 inputs under bf16-mixed autocast (trainer.py:450) with D=384 take the fused path."""
    import json

    z = np.load(golden_dir / "losses_anchor_m64_d384.npz")
    cand = xr.PoolCandidates(torch.from_numpy(z["pos"]).cuda(), torch.from_numpy(z["neg"]).cuda())
    q = torch.from_numpy(z["query"]).cuda().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = xr.InfoNCELoss(xr.LossConfig())(q, cand)
    want, _, _, _ = orc.lean_loss("InfoNCELoss", z["query"], z["pos"], z["neg"], orc.Config(),
                                  with_grad=True, logits_dtype="bf16")
    assert float(loss) == pytest.approx(want, rel=2e-3)


@pytest.mark.parametrize("u,n,k", [(1, 5000, 20), (6, 70001, 100), (130, 20000, 100), (257, 3001, 50)])
def test_groupmax_retrieval_equals_full_scan(xr, u, n, k):
    """tcgen05 group-max path vs the stable-sort oracle on exact-arithmetic inputs (small integers:
    every bf16 product / fp32 sum is exact, so indices must be IDENTICAL, ties included) and vs the
    library's own unfused path."""
    rng = np.random.default_rng(n)
    cat = rng.integers(-2, 3, size=(n, 384)).astype(np.float32)
    cat[5] = cat[3]                      # duplicate rows: ties -> lower id first
    qs = rng.integers(-2, 3, size=(u, 384)).astype(np.float32)
    excl = [list(rng.integers(0, n, size=int(rng.integers(0, 60)))) for _ in range(u)]
    cfg = dict(index_metric="dot", dtype="bf16")
    fused = xr.index.ExactIndex(xr.index.ExactIndexConfig(**cfg)).set_catalog(torch.from_numpy(cat).cuda())
    plain = xr.index.ExactIndex(xr.index.ExactIndexConfig(fused=False, **cfg)).set_catalog(torch.from_numpy(cat).cuda())
    q = torch.from_numpy(qs).cuda()
    s1, i1 = fused.search_batch(q, excl, k)
    s2, i2 = plain.search_batch(q, excl, k)
    want_s, want_i = orc.exact_search(qs, cat, k, excl, metric="dot")
    assert np.array_equal(i1.cpu().numpy(), want_i)
    assert np.array_equal(s1.cpu().numpy(), want_s)
    assert torch.equal(i1, i2) and torch.equal(s1, s2)


def test_groupmax_retrieval_random_cosine(xr):
    """Random bf16 catalog, cosine metric: same ids as the unfused path wherever neighbouring
    scores are separated by more than fp32 accumulation-order noise; recall@100 = 1 vs fp64."""
    rng = np.random.default_rng(1)
    n, u, k = 200_000, 64, 100
    cat = rng.standard_normal((n, 384)).astype(np.float32)
    qs = rng.standard_normal((u, 384)).astype(np.float32)
    fused = xr.index.ExactIndex().set_catalog(torch.from_numpy(cat).cuda())
    plain = xr.index.ExactIndex(xr.index.ExactIndexConfig(fused=False)).set_catalog(torch.from_numpy(cat).cuda())
    q = torch.from_numpy(qs).cuda()
    s1, i1 = fused.search_batch(q, None, k)
    s2, i2 = plain.search_batch(q, None, k)
    assert torch.allclose(s1, s2, atol=2e-6)
    same = (i1 == i2).float().mean().item()
    assert same > 0.995, same
    # the arithmetic the index defines: dot products of the normalised-then-bf16-rounded rows
    catn = fused.catalog.float().cpu().numpy()
    qn, _ = xr.ops.normalize_rows(q, 1e-12, torch.bfloat16)
    want_s, want_i = orc.exact_search(qn.float().cpu().numpy(), catn, k, None, metric="dot", dtype=np.float64)
    recall = np.mean([len(set(i1[r].tolist()) & set(want_i[r].tolist())) / k for r in range(u)])
    assert recall >= 0.999, recall


@pytest.mark.parametrize("u,n", [(129, 1000), (256, 12800), (300, 4097), (1000, 130), (513, 64), (260, 70000)])
def test_groupmax_cta_pair_kernel_matches_reference(xr, u, n):
    """u > 128 runs on CTA pairs (tcgen05 cta_group::2, 256 queries per pair); group maxima must
    equal the maxima of the fp32-accumulated scores, in NATURAL order (column c = rows [16c, 16c+16),
    so that top-k ties between groups resolve towards the lower row), and agree with the single-CTA
    variant."""
    from xfmr_rec_b200 import _native as N, ops

    g = torch.Generator(device="cuda").manual_seed(u * 7 + n)
    q = torch.randn((u, 384), generator=g, device="cuda").bfloat16()
    cat = torch.randn((n, 384), generator=g, device="cuda").bfloat16()
    ng = (n + 15) // 16
    want = torch.full((u, ng * 16), float("-inf"), device="cuda")
    want[:, :n] = q.float() @ cat.float().T
    want = want.view(u, ng, 16).amax(-1)                       # natural order: group g = rows [16g, 16g+16)
    lib = N.lib()
    got_pair = ops.score_groupmax(q, cat)
    assert got_pair.size(1) >= ng
    torch.testing.assert_close(got_pair[:, :ng], want, rtol=2e-3, atol=2e-3)
    assert bool((got_pair[:, ng:] == float("-inf")).all())     # columns past the catalog: never selected
    lib.xr_fused_wait_stats(4, None)          # profiling switch: force the single-CTA kernel
    try:
        got_single = ops.score_groupmax(q, cat)[:, :ng].clone()
    finally:
        lib.xr_fused_wait_stats(0, None)
    torch.testing.assert_close(got_single, want, rtol=2e-3, atol=2e-3)
    assert torch.equal(got_single, got_pair[:, :ng])           # same MMA arithmetic in both kernels


@pytest.mark.parametrize("u,n,stride", [(5, 70000, 7), (64, 4097, 3), (200, 70000, 5), (300, 200000, 32)])
def test_groupmax_sampled_tiles(xr, u, n, stride):
    """tile_stride = s: only every s-th tile (64 rows for U <= 128, 128 above) is scored; column
    (T/16) t + g holds rows [T t s + 16 g, +16).  The sample's maxima are maxima of real catalog rows,
    which is all the threshold logic needs."""
    from xfmr_rec_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(u + n)
    q = torch.randn((u, 384), generator=g, device="cuda").bfloat16()
    cat = torch.randn((n, 384), generator=g, device="cuda").bfloat16()
    T = 128 if u > 128 else 64
    nt = ((n + T - 1) // T + stride - 1) // stride
    full = torch.full((u, (n + T * stride) // 16 * 16 + T * stride), float("-inf"), device="cuda")
    full[:, :n] = q.float() @ cat.float().T
    rows = (torch.arange(nt, device="cuda")[:, None] * (T * stride) + torch.arange(T, device="cuda")[None]).reshape(-1)
    want = full[:, rows].view(u, nt * T // 16, 16).amax(-1)
    got = ops.score_groupmax(q, cat, stride)
    assert got.size(1) >= want.size(1)
    torch.testing.assert_close(got[:, :want.size(1)], want, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("cap_b", [None, 8])
@pytest.mark.parametrize("u,n", [(3, 9000), (128, 30001), (129, 30001), (400, 12800)])
def test_score_filter_survivors(xr, u, n, cap_b):
    """xr_score_filter: the survivors of a query are EXACTLY the rows whose score is >= its threshold
    (exact-arithmetic inputs), wherever the scoring kernel stored them: sub-buckets filled without atomics
    by their owning lanes, and -- with cap_b = 8 slots, which most sub-buckets overrun -- the query's
    overflow list.  The finalize step turns them into the oracle's top-k."""
    from xfmr_rec_b200 import ops

    rng = np.random.default_rng(u * 31 + n)
    cat = rng.integers(-2, 3, size=(n, 384)).astype(np.float32)
    qs = rng.integers(-2, 3, size=(u, 384)).astype(np.float32)
    scores = qs @ cat.T
    th = np.sort(scores, axis=1)[:, -150].copy()               # ~150+ survivors per query (ties included)
    if n <= 16384:
        th[0] = -np.inf                                         # a query that keeps everything
    th[-1] = np.inf                                             # and one that keeps nothing
    q16, c16 = torch.from_numpy(qs).cuda().bfloat16(), torch.from_numpy(cat).cuda().bfloat16()
    tht = torch.from_numpy(th).cuda()
    fs = ops.score_filter(q16, c16, tht, expected_survivors=n if n <= 16384 else 600, cap_b=cap_b, ovf_cap=16384)
    counts = fs.counts().cpu().numpy()
    for r, (sc, ro) in enumerate(fs.lists()):
        want = np.nonzero(scores[r] >= th[r])[0]
        assert counts[r] == len(want), (r, counts[r], len(want))
        order = np.argsort(ro)
        assert np.array_equal(ro[order], want)
        assert np.array_equal(sc[order], scores[r, want])
    if cap_b is not None and n <= 16384:
        assert int(fs.o_count.max()) > 0                        # the keep-everything query spilled: overflow tier exercised
    k = 20
    s, i, flags = ops.filter_finalize(fs, n, tht, 60, k, row_offset=1000)
    # the last query kept nothing: it cannot vouch for the rows below its (infinite) threshold -> flag 4
    assert int(flags.item()) == 4
    want_s, want_i = orc.exact_search(qs[:-1], cat, k, None, metric="dot")
    assert np.array_equal(i.cpu().numpy()[:-1], want_i + 1000)
    assert np.array_equal(s.cpu().numpy()[:-1], want_s)
    assert bool((i[-1] == -1).all())


@pytest.mark.parametrize("u", [64, 256])
@pytest.mark.parametrize("n,case", [(20000, "identical"), (50000, "identical"), (120000, "dups")])
def test_exact_index_ties_resolve_to_lowest_rows(xr, u, n, case):
    """Heavy ties on the bf16 tensor-core path, U <= 128 (single-CTA kernel) and U > 128 (CTA pairs):
    a catalog of identical rows, and 4,096 scattered duplicates of the best row.  The result must be
    the LOWEST row ids (north_star: ties broken by lower item id).  20,000 identical rows all survive
    the filter; 50,000 overflow the survivor lists and take the materialised path; both are exact."""
    rng = np.random.default_rng(n + u)
    k = 100
    qs = rng.standard_normal((u, 384)).astype(np.float32)
    if case == "identical":
        cat = np.tile(rng.standard_normal((1, 384)).astype(np.float32), (n, 1))
        want = np.tile(np.arange(k), (u, 1))
    else:
        cat = rng.standard_normal((n, 384)).astype(np.float32) * 0.1
        dup_rows = np.sort(rng.choice(n, size=4096, replace=False))
        cat[dup_rows] = qs.mean(0, keepdims=True) * 50.0        # far better than any random row, for every query
        want = np.tile(dup_rows[:k], (u, 1))
        ok = 50.0 * (qs @ qs.mean(0)) > 15.0                     # queries for which the duplicates are the best rows
        assert ok.mean() > 0.5
    idx = xr.index.ExactIndex(xr.index.ExactIndexConfig(index_metric="dot", dtype="bf16")).set_catalog(
        torch.from_numpy(cat).cuda())
    s, i = idx.search_batch(torch.from_numpy(qs).cuda(), None, k)
    got = i.cpu().numpy()
    if case == "identical":
        assert np.array_equal(got, want)
    else:
        assert np.array_equal(got[ok], want[ok])
    plan = idx.compile_search(u, k)
    ps, pi = plan(torch.from_numpy(qs).cuda())
    assert torch.equal(pi, i) and torch.equal(ps, s)


def _cfg(ops, logits_bf16=False, **kw):
    class C:
        pass

    c = C()
    for k, v in {**dict(mask_false_negatives=True, num_hard_negatives=0, scale=1.0, margin=0.5), **kw}.items():
        setattr(c, k, v)
    return ops.make_cfg(c, logits_bf16=logits_bf16)


DOT_SLOTS = ["InfoNCELoss", "NCELoss", "PairwiseHingeLoss", "PairwiseLogisticLoss"]
COS_SLOTS = ["AlignmentLoss", "ContrastiveLoss", "AlignmentContrastiveLoss"]


@pytest.mark.parametrize("m,cn", [(1, 1), (128, 64), (130, 65), (301, 777), (700, 3000), (1500, 4200)])
def test_fused_all_losses_and_stats_one_pass(xr, m, cn):
    """xr_fused_pool_all: every loss of a logit family + the LogitsStatistics block from ONE
    tensor-core pass (trainer.py:250-263) vs the oracle on the bf16 operands the kernel saw, and
    vs the library's own materialised path (xr_logits_pool + xr_rowloss)."""
    from xfmr_rec_b200 import _native as N, ops

    q, pos, neg = make_inputs(m, cn, seed=m * 3 + cn)
    for cfg_kw, lbf in [({}, True), ({"scale": 8.0, "margin": 0.2}, True), ({"mask_false_negatives": False}, False)]:
        for cosine in (False, True):
            lb = lbf and not cosine
            cfg = _cfg(ops, logits_bf16=lb, **cfg_kw)
            qt, pt, nt = bf(q), bf(pos), bf(neg)
            if cosine:
                qt, pt, nt = (ops.normalize_rows(torch.from_numpy(x).cuda(), 1e-8, torch.bfloat16)[0]
                              for x in (q, pos, neg))
            losses, stats = ops.fused_pool_all(qt, pt, nt, cfg, cosine)
            # the materialised path of the same library on the same operands
            logits = ops.logits_pool(qt, pt, nt)
            l2, s2, _ = ops.rowloss(logits, cn + 1, cfg, N.TARGET_LAST, None, -1, want_stats=True)
            losses, stats, l2, s2 = (t.cpu().numpy() for t in (losses, stats, l2, s2))
            names = COS_SLOTS if cosine else DOT_SLOTS
            for name in names:
                k = N.LOSS_KIND[name]
                assert losses[k] == pytest.approx(l2[k], rel=2e-3, abs=2e-3), (name, cfg_kw, cosine)
            # oracle on the bf16 operands
            qh, ph, nh = (t.float().cpu().numpy() for t in (qt, pt, nt))
            ocfg = orc.Config(**cfg_kw)
            lg = orc.lean_logits(qh, ph, nh)
            lg = (orc.round_bf16(lg.astype(np.float32)) if lb else lg.astype(np.float32)).astype(np.float64)
            tgt = np.zeros(m, np.int64)
            mask = orc.mask_false_negatives(lg, tgt, ocfg)
            for name in names:
                want = orc.loss_from_logits(name, lg, tgt, mask, ocfg, with_grad=False)
                assert losses[N.LOSS_KIND[name]] == pytest.approx(want, rel=2e-3, abs=2e-3), (name, cfg_kw)
            # statistics block: density, rows, pos {sum, sumsq, min, max}, neg {count, sum, sumsq, min, max}
            assert stats[1] == m and stats[11] == cn
            # counts depend on '<' decisions at fp32-accumulation-order resolution: allow a few flips
            assert abs(stats[6] - s2[6]) <= max(2.0, 2e-4 * s2[6]), (stats[6], s2[6])
            np.testing.assert_allclose(stats[[0, 2, 3]], s2[[0, 2, 3]], rtol=2e-3, atol=2e-2)
            np.testing.assert_allclose(stats[[4, 5]], s2[[4, 5]], rtol=1e-3, atol=1e-3)
            if cosine:   # negative-logit statistics exist for the dot family only (losses.py:383-386)
                continue
            np.testing.assert_allclose(stats[[7, 8]], s2[[7, 8]], rtol=2e-3, atol=2e-2)
            if s2[6] > 0:
                np.testing.assert_allclose(stats[[9, 10]], s2[[9, 10]], rtol=1e-3, atol=2e-3)


def test_evaluate_all_matches_modules(xr):
    """evaluate_all (= compute_losses, trainer.py:213-264): same numbers as the seven loss modules
    and LogitsStatistics called one by one.  With a gradient wanted it is ONE tensor-core pass
    (xr_fused_pool_loss_mon: train loss + dq + both families + statistics); without, one forward pass per
    family -- the dot family and the statistics agree between the two to fp32 summation order, the cosine
    family to bf16 tolerance; loss value and gradient are those of the InfoNCE module bit for bit."""
    q, pos, neg = make_inputs(300, 1000, seed=11)
    cfg = xr.LossConfig()
    qt = bf(q).requires_grad_(True)
    cand = xr.PoolCandidates(bf(pos), bf(neg))
    out, stats = xr.losses.evaluate_all(cfg, qt, cand)
    out["loss/InfoNCELoss"].backward()
    assert qt.grad is not None and bool(torch.isfinite(qt.grad).all())
    q2 = bf(q).requires_grad_(True)
    l2 = xr.InfoNCELoss(cfg)(q2, cand)
    l2.backward()
    assert torch.equal(out["loss/InfoNCELoss"].detach(), l2.detach()) and torch.equal(qt.grad, q2.grad)
    with torch.no_grad():
        out_ng, stats_ng = xr.losses.evaluate_all(cfg, bf(q), cand)
    for k, v in out_ng.items():
        tol = 4e-3 if k.split("/")[1] in orc.COSINE_LOSSES else 1e-6
        assert float(out[k]) == pytest.approx(float(v), rel=tol, abs=tol), k
    for k, v in stats_ng.items():
        assert stats[k] == pytest.approx(v, rel=1e-6, abs=1e-9), k
    for name in orc.LOSS_NAMES:
        want = float(getattr(xr, name)(cfg)(bf(q), cand))
        assert float(out[f"loss/{name}"]) == pytest.approx(want, rel=2e-3, abs=2e-3), name
    lb = orc.round_bf16(orc.lean_logits(orc.round_bf16(q), orc.round_bf16(pos), orc.round_bf16(neg)).astype(np.float32)).astype(np.float64)
    tgt = np.zeros(q.shape[0], np.int64)
    want_stats = orc.logits_statistics(lb, tgt, orc.mask_false_negatives(lb, tgt, orc.Config()), orc.Config())
    for k, v in want_stats.items():
        assert stats[k] == pytest.approx(v, rel=5e-3, abs=5e-3), k


@pytest.mark.parametrize("u,n,k", [(5, 40000, 20), (140, 30000, 50)])
def test_groupmax_two_phase_rescore_with_adversarial_exclusions(xr, u, n, k):
    """Exclusion lists: phase 1 re-scores top_k + 28 groups, phase 2 the conservative remainder only
    for queries whose k-th surviving score does not clear the next group maximum.  Adversarial case:
    the excluded rows ARE the best-scoring rows (a user's history scores high), so phase 2 must run;
    exact-arithmetic inputs -> indices identical to the stable-sort oracle, ties included."""
    rng = np.random.default_rng(n + u)
    cat = rng.integers(-2, 3, size=(n, 384)).astype(np.float32)
    cat[7] = cat[3]
    qs = rng.integers(-2, 3, size=(u, 384)).astype(np.float32)
    _, top = orc.exact_search(qs, cat, 160, None, metric="dot")
    excl = []
    for r in range(u):
        if r % 3 == 0:
            excl.append([int(x) for x in top[r, :150]])               # everything phase 1 would find
        elif r % 3 == 1:
            excl.append([int(x) for x in top[r, ::2][:60]])           # every other top row
        else:
            excl.append([int(x) for x in rng.integers(0, n, size=30)])  # random: phase 1 suffices
    idx = xr.index.ExactIndex(xr.index.ExactIndexConfig(index_metric="dot", dtype="bf16")).set_catalog(
        torch.from_numpy(cat).cuda())
    q = torch.from_numpy(qs).cuda()
    s, i = idx.search_batch(q, excl, k)
    want_s, want_i = orc.exact_search(qs, cat, k, excl, metric="dot")
    assert np.array_equal(i.cpu().numpy(), want_i)
    assert np.array_equal(s.cpu().numpy(), want_s)
    plan = idx.compile_search(u, k, max_exclusions=150)
    ps, pi = plan(q, xr.ops._csr(excl, q.device))
    assert torch.equal(pi, i) and torch.equal(ps, s)


def test_fused_under_autocast_with_margin_and_scale_vs_the_reference(xr, golden_dir):
    """The reference under bf16 autocast with scale = 5, margin = 0.3 (second fixture of
    make_golden_autocast_d384.py): there it also rounds target * (1 - margin), logits - that and logits * scale
    to bf16.  The kernels emulate the scale rounding and keep the margin arithmetic in fp32: losses
    agree with the reference's own autocast values to 5e-4 (hinge: 2e-4 measured, the others 1e-6), and the
    gradient is an order of magnitude closer to the float64 oracle than the reference's autocast gradient
    (whose bf16 differences flip hinge indicators: 3.3e-2 norm-wise)."""
    from xfmr_rec_b200 import _native as N, ops

    z = np.load(golden_dir / "losses_pool_autocast_bf16_d384_margin.npz")
    q, p, n = z["query"], z["pos"], z["neg"]
    kw = dict(scale=float(z["cfg"][0]), margin=float(z["cfg"][1]))
    for name in ("InfoNCELoss", "NCELoss", "PairwiseHingeLoss", "PairwiseLogisticLoss"):
        want, want_dq, _, _ = orc.lean_loss(name, q, p, n, orc.Config(**kw), with_grad=True, logits_dtype="bf16")
        loss, dq, _ = ops.fused_pool_loss(bf(q), bf(p), bf(n), N.LOSS_KIND[name],
                                          ops.make_cfg(xr.LossConfig(**kw), logits_bf16=True))
        loss = float(loss.view(torch.float32)[2])
        assert loss == pytest.approx(float(z[f"autocast/loss/{name}"]), rel=5e-4), name
        dq = dq.cpu().numpy().astype(np.float64)
        ref = z[f"autocast/dq/{name}"].astype(np.float64)
        e_ours = np.linalg.norm(dq - want_dq) / np.linalg.norm(want_dq)
        e_ref = np.linalg.norm(ref - want_dq) / np.linalg.norm(want_dq)
        assert e_ours <= 1e-3 and e_ours < 0.5 * e_ref, (name, e_ours, e_ref)


def test_fused_bf16_gradient_vs_the_references_own_autocast(xr, golden_dir):
    """The reference under bf16-mixed autocast (trainer.py:450) executed at D = 384 (tests/golden/
    make_golden_autocast_d384.py): its losses are reproduced to 1e-6, and the fused kernel's fp32 gradient
    is CLOSER to the float64 oracle (same bf16-rounded operands and logits) than the reference's own
    autocast gradient, whose bmm backward rounds dlogits and dq to bf16."""
    from xfmr_rec_b200 import _native as N, ops

    z = np.load(golden_dir / "losses_pool_autocast_bf16_d384.npz")
    q, p, n = z["query"], z["pos"], z["neg"]
    for name in ("InfoNCELoss", "NCELoss", "PairwiseHingeLoss", "PairwiseLogisticLoss"):
        want, want_dq, _, _ = orc.lean_loss(name, q, p, n, orc.Config(), with_grad=True, logits_dtype="bf16")
        loss, dq, _ = ops.fused_pool_loss(bf(q), bf(p), bf(n), N.LOSS_KIND[name],
                                          ops.make_cfg(xr.LossConfig(), logits_bf16=True))
        loss = float(loss.view(torch.float32)[2])
        assert loss == pytest.approx(float(z[f"autocast/loss/{name}"]), rel=5e-5), name
        assert loss == pytest.approx(want, rel=1e-6), name
        dq = dq.cpu().numpy().astype(np.float64)
        ref = z[f"autocast/dq/{name}"].astype(np.float64)
        e_ours = np.linalg.norm(dq - want_dq) / np.linalg.norm(want_dq)
        e_ref = np.linalg.norm(ref - want_dq) / np.linalg.norm(want_dq)
        assert e_ours <= 1e-3 and e_ours < e_ref, (name, e_ours, e_ref)
        assert np.linalg.norm(dq - ref) <= 4e-3 * np.linalg.norm(ref), name


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("XR_SWEEP_SEEDS", "24"))))
def test_search_random_shapes_against_the_oracle(xr, seed):
    """Differential sweep of the tensor-core search (xr_score_topk through ExactIndex.search_batch and a
    compiled SearchPlan): random catalog sizes (not multiples of any tile), query counts on both sides of the
    128-query kernel switch, k, exclusion lists with duplicates, out-of-range and repeated ids, empty lists,
    catalogs with many exactly tied rows.  Integer-valued data make every score exact in bf16/fp32, so ids and
    scores must equal the stable-sort oracle bit for bit -- whether the filter path answers or a device flag
    routes the batch to the materialised search."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.choice([1, 17, 129, 1000, 4097, 20011, 65537, 131072 + 5]))
    u = int(rng.choice([1, 2, 31, 128, 129, 200, 257]))
    k = int(min(n, rng.choice([1, 5, 50, 100, 128])))
    lo, hi = (-1, 2) if seed % 3 == 0 else (-3, 4)          # narrow range: heavy ties
    cat = rng.integers(lo, hi, size=(n, 384)).astype(np.float32)
    if n > 40:
        cat[rng.integers(0, n, size=20)] = cat[0]             # exact duplicates of one row
    qs = rng.integers(lo, hi, size=(u, 384)).astype(np.float32)
    max_excl = int(rng.choice([0, 3, 40, 200]))
    excl = []
    for r in range(u):
        m_e = int(rng.integers(0, max_excl + 1))
        ids = rng.integers(-2, n + 3, size=m_e)               # some ids are out of range: ignored
        if m_e > 2:
            ids[1] = ids[0]                                   # a repeated id
        excl.append([int(x) for x in ids])
    metric = "dot" if seed % 2 == 0 else "cosine"
    if metric == "cosine":
        cat[np.abs(cat).sum(1) == 0] = 1.0                    # no zero-norm rows: cosine of integer rows is not exact,
    idx = xr.index.ExactIndex(xr.index.ExactIndexConfig(index_metric="dot", dtype="bf16")).set_catalog(
        torch.from_numpy(cat).cuda())                         # ... so the sweep scores dot products either way
    q = torch.from_numpy(qs).cuda()
    want_s, want_i = orc.exact_search(qs, cat, k, excl if max_excl else None, metric="dot")
    want_i = np.where(np.isneginf(want_s), -1, want_i)       # filtered rows are never returned (index.py:246): id -1
    s, i = idx.search_batch(q, excl if max_excl else None, k, max_exclusions=max_excl or None)
    assert np.array_equal(i.cpu().numpy(), want_i), (n, u, k, max_excl)
    assert np.array_equal(s.cpu().numpy(), want_s)
    plan = idx.compile_search(u, k, max_exclusions=max_excl)
    ps, pi = plan(q, xr.ops._csr(excl, q.device) if max_excl else None)
    assert torch.equal(pi, i) and torch.equal(ps, s)


@pytest.mark.parametrize("seed", range(16))
def test_fused_random_shapes_against_the_oracle(xr, seed):
    """Stream-K edge cases of the fused loss kernel: random (rows, pool) shapes with extreme aspect ratios --
    one candidate tile and thousands of rows (every CTA crosses many row blocks: many segments), one row block
    and a long pool (one row block split over all CTAs), fewer tiles than CTAs, sizes one off every tile
    boundary -- for a random fused kind and config; loss and dL/dq against the float64 oracle on the operands
    the kernel saw, and forward-only == forward+backward."""
    rng = np.random.default_rng(500 + seed)
    shapes = [(4000, 64), (5000, 1), (1, 9000), (128, 20000), (129, 12801), (17, 17), (2500, 130), (640, 641),
              (3000, 700), (127, 63)]
    m, cn = shapes[seed % len(shapes)]
    if seed >= len(shapes):
        m, cn = int(rng.integers(1, 3000)), int(rng.integers(1, 3000))
    name = FUSED[int(rng.integers(0, len(FUSED)))]
    cfg_kw, lbf = [({}, True), ({"scale": 5.0, "margin": 0.3}, True), ({"mask_false_negatives": False}, False),
                   ({"scale": 0.5}, False)][int(rng.integers(0, 4))]
    q, pos, neg = make_inputs(m, cn, seed=seed)
    loss, dq, (qh, ph, nh, q_inv) = run_fused(xr, name, q, pos, neg, cfg_kw, lbf)
    want, want_dq = oracle_on_bf16(name, qh, ph, nh, q_inv, cfg_kw, lbf)
    assert loss == pytest.approx(want, rel=2e-3, abs=2e-3), (name, m, cn, cfg_kw, loss, want)
    assert np.linalg.norm(dq - want_dq) <= 2e-3 * np.linalg.norm(want_dq) + 1e-6, (name, m, cn, cfg_kw)
    loss_f, _, _ = run_fused(xr, name, q, pos, neg, cfg_kw, lbf, want_grad=False)
    assert loss_f == loss
