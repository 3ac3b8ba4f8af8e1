"""GPU parity tests for the sync-free scoring-and-loss step (xr_pool_step / PoolLossStep):
bit-identical to compute_embeds + loss module + backward (the same kernels, planned on the device
instead of the host), and within tolerance of the oracle."""

import numpy as np
import pytest
import torch

from oracle import xfmr_oracle as orc

pytestmark = pytest.mark.gpu

KINDS = ["InfoNCELoss", "NCELoss", "PairwiseHingeLoss", "PairwiseLogisticLoss",
         "ContrastiveLoss", "AlignmentContrastiveLoss"]   # dot family + the cosine (CCL) family


@pytest.fixture(scope="module")
def xr():
    import xfmr_rec_b200 as pkg

    if not torch.cuda.is_available() or torch.cuda.get_device_capability()[0] != 10:
        pytest.skip("needs an sm_100 device")
    assert pkg._native.lib().xr_fused_available() & 1, "tcgen05 kernels missing from the build"
    return pkg


def modular(xr, emb, loss_fn, b, tok_dtype):
    """The drop-in path: compute_embeds -> loss -> backward (one host sync for the counts)."""
    tok = torch.from_numpy(b["token_embeddings"]).cuda().to(tok_dtype).requires_grad_(True)
    hist, pos, neg = (torch.from_numpy(b[k]).cuda() for k in
                      ("history_item_idx", "pos_item_idx", "neg_item_idx"))
    out = xr.models.compute_embeds(emb, tok, hist, pos, neg, candidate_dtype=torch.bfloat16)
    q = out["query_embed"]
    if q.dtype != torch.bfloat16:   # the step rounds fp32 encoder output to bf16 operands
        q = q.bfloat16()
    loss = loss_fn(q, out["candidate_embed"])
    loss.backward()
    return loss.detach(), tok.grad, int(out["attention_mask"].sum()), q.size(0)


def run_step(step, b, dtype):
    tok = torch.from_numpy(b["token_embeddings"]).to(dtype)
    return step(tok.pin_memory(), torch.from_numpy(b["history_item_idx"]).pin_memory(),
                torch.from_numpy(b["pos_item_idx"]).pin_memory(),
                torch.from_numpy(b["neg_item_idx"]).pin_memory())


@pytest.mark.parametrize("name", KINDS)
@pytest.mark.parametrize("graph", [True, False])
def test_step_matches_modular_path_bitwise(xr, name, graph):
    b = orc.synth_batch(3000, 16, 60, dim=384, seed=3)
    emb = xr.models.ItemEmbeddings(torch.from_numpy(b["table"]), add_padding_row=False).cuda()
    loss_fn = getattr(xr, name)(xr.LossConfig())
    want_loss, want_grad, m_a, m = modular(xr, emb, loss_fn, b, torch.bfloat16)
    step = xr.PoolLossStep(emb, loss_fn, 16, 60, use_graph=graph)
    loss, dtok = run_step(step, b, torch.bfloat16)
    torch.cuda.synchronize()
    assert step.row_counts() == (m_a, m)
    assert torch.equal(loss, want_loss), (float(loss), float(want_loss))
    assert torch.equal(dtok.reshape(want_grad.shape), want_grad)


@pytest.mark.parametrize("cfg_kw", [dict(mask_false_negatives=False), dict(scale=20.0, margin=0.2),
                                    dict(mask_false_negatives=False, scale=5.0)])
def test_step_config_variants(xr, cfg_kw):
    b = orc.synth_batch(2000, 8, 100, dim=384, seed=5)
    emb = xr.models.ItemEmbeddings(torch.from_numpy(b["table"]), add_padding_row=False).cuda()
    for name in ("InfoNCELoss", "PairwiseLogisticLoss"):
        loss_fn = getattr(xr, name)(xr.LossConfig(**cfg_kw))
        want_loss, want_grad, _, _ = modular(xr, emb, loss_fn, b, torch.bfloat16)
        loss, dtok = run_step(xr.PoolLossStep(emb, loss_fn, 8, 100), b, torch.bfloat16)
        assert torch.equal(loss, want_loss)
        assert torch.equal(dtok.reshape(want_grad.shape), want_grad)


def test_step_reuse_with_shrinking_batches(xr):
    """One step object, consecutive batches with fewer and fewer valid rows: rows left over from
    an earlier, larger batch must never leak into a later result (zero fill of the last tile)."""
    table = orc.synth_batch(1500, 1, 1, seed=0)["table"]
    emb = xr.models.ItemEmbeddings(torch.from_numpy(table), add_padding_row=False).cuda()
    loss_fn = xr.InfoNCELoss(xr.LossConfig())
    step = xr.PoolLossStep(emb, loss_fn, 12, 70)
    for seed, frac in ((1, 0.0), (2, 0.5), (3, 0.9), (4, 0.2)):
        b = orc.synth_batch(1500, 12, 70, dim=384, seed=seed, pos_pad_frac=frac, table=table)
        if seed == 3:   # also shorten the histories a lot
            b["history_item_idx"][:, 9:] = 0
        want_loss, want_grad, m_a, m = modular(xr, emb, loss_fn, b, torch.bfloat16)
        loss, dtok = run_step(step, b, torch.bfloat16)
        torch.cuda.synchronize()
        assert step.row_counts() == (m_a, m)
        assert torch.equal(loss, want_loss)
        assert torch.equal(dtok.reshape(want_grad.shape), want_grad)
        assert bool(torch.isfinite(dtok.float()).all())


def test_step_empty_batch(xr):
    b = orc.synth_batch(500, 4, 20, dim=384, seed=9)
    b["pos_item_idx"][:] = 0          # no position has a positive: M = 0
    emb = xr.models.ItemEmbeddings(torch.from_numpy(b["table"]), add_padding_row=False).cuda()
    step = xr.PoolLossStep(emb, xr.InfoNCELoss(xr.LossConfig()), 4, 20)
    loss, dtok = run_step(step, b, torch.bfloat16)
    torch.cuda.synchronize()
    assert step.row_counts()[1] == 0
    assert float(loss) == 0.0 and not bool(dtok.any())


def test_step_fp32_tokens_and_oracle(xr):
    """fp32 encoder output: operands rounded to bf16 inside the gather; loss and gradient within
    the bf16 tolerance (2e-3) of the oracle run on bf16-rounded operands and bf16 logits."""
    b = orc.synth_batch(800, 6, 40, dim=384, seed=11)
    emb = xr.models.ItemEmbeddings(torch.from_numpy(b["table"]), add_padding_row=False).cuda()
    loss_fn = xr.InfoNCELoss(xr.LossConfig())
    step = xr.PoolLossStep(emb, loss_fn, 6, 40, token_dtype=torch.float32, logits_bf16=True)
    loss, dtok = run_step(step, b, torch.float32)
    want = orc.compute_embeds(b["table"], b["token_embeddings"], b["history_item_idx"],
                              b["pos_item_idx"], b["neg_item_idx"], dense=False)
    ref, dq, _, _ = orc.lean_loss("InfoNCELoss", orc.round_bf16(want["query_embed"]),
                                  orc.round_bf16(want["pos_embed"]), orc.round_bf16(want["neg_embed"]),
                                  orc.Config(), with_grad=True, logits_dtype="bf16")
    assert abs(float(loss) - ref) <= 2e-3 * abs(ref)
    got = dtok.reshape(-1, 384).float().cpu().numpy()
    mask = (b["history_item_idx"].reshape(-1) != 0) & (b["pos_item_idx"].reshape(-1) != 0)
    assert not got[~mask].any()
    err = np.abs(got[mask] - dq).max()
    assert err <= 2e-3 * np.abs(dq).max(), err


def test_step_rejects_what_it_does_not_serve(xr):
    b = orc.synth_batch(100, 2, 8, dim=384, seed=1)
    emb = xr.models.ItemEmbeddings(torch.from_numpy(b["table"]), add_padding_row=False).cuda()
    with pytest.raises(NotImplementedError):
        xr.PoolLossStep(emb, xr.AlignmentLoss(xr.LossConfig()), 2, 8)     # no negatives: nothing to fuse
    with pytest.raises(NotImplementedError):
        xr.PoolLossStep(emb, xr.InfoNCELoss(xr.LossConfig(num_hard_negatives=5)), 2, 8)
    with pytest.raises(xr._native.NativeError):
        xr.PoolLossStep(xr.models.ItemEmbeddings(torch.from_numpy(b["table"]), add_padding_row=False),
                        xr.InfoNCELoss(xr.LossConfig()), 2, 8)


@pytest.mark.parametrize("one_pass", [True, False])
@pytest.mark.parametrize("graph", [True, False])
@pytest.mark.parametrize("cfg_kw", [{}, dict(mask_false_negatives=False, scale=4.0, margin=0.2)])
def test_step_with_monitor_equals_evaluate_all(xr, graph, cfg_kw, one_pass):
    """PoolLossStep(monitor=True): the train loss + gradient AND everything compute_losses logs
    (trainer.py:250-263: LogitsStatistics + all seven losses) from one sync-free sequence; same
    numbers as evaluate_all on the module path.  one_pass (default): ONE tensor-core pass -- the train
    kernel accumulates the monitoring sums too; train loss and gradient stay bit-identical, the dot family
    and the statistics equal the separate all-losses pass (same arithmetic on the same scores), the cosine
    family is evaluated as (score / |q|) / |n| with fp32 inverse norms and agrees to bf16 tolerance with
    the pass over bf16-normalised operands.  one_pass=False: the three-pass sequence, identical kernels."""
    b = orc.synth_batch(3000, 16, 60, dim=384, seed=4)
    emb = xr.models.ItemEmbeddings(torch.from_numpy(b["table"]), add_padding_row=False).cuda()
    cfg = xr.LossConfig(**cfg_kw)
    loss_fn = xr.InfoNCELoss(cfg)
    step = xr.PoolLossStep(emb, loss_fn, 16, 60, use_graph=graph, monitor=True, monitor_one_pass=one_pass)
    assert step.monitor_one_pass == one_pass
    for rep in range(2):     # replay twice: no state may leak between steps
        loss, dtok = run_step(step, b, torch.bfloat16)
        got, got_stats = step.loss_dict()
        tok = torch.from_numpy(b["token_embeddings"]).cuda().bfloat16().requires_grad_(True)
        hist, pos, neg = (torch.from_numpy(b[k]).cuda() for k in
                          ("history_item_idx", "pos_item_idx", "neg_item_idx"))
        out = xr.models.compute_embeds(emb, tok, hist, pos, neg, candidate_dtype=torch.bfloat16)
        want, want_stats = xr.losses.evaluate_all(cfg, out["query_embed"], out["candidate_embed"])
        want["loss/InfoNCELoss"].backward()
        assert torch.equal(loss, want["loss/InfoNCELoss"].detach())
        assert torch.equal(dtok.reshape(tok.grad.shape), tok.grad)
        for k, v in want.items():
            cos_key = k.split("/")[1] in orc.COSINE_LOSSES
            # (evaluate_all on the module path is one-pass too: the three-pass step differs in the cosine family)
            tol = 4e-3 if (cos_key and not one_pass) else 1e-6
            assert float(got[k]) == pytest.approx(float(v), rel=tol, abs=tol), k
        assert got_stats.keys() == want_stats.keys()
        for k, v in want_stats.items():
            assert got_stats[k] == pytest.approx(v, rel=1e-6, abs=1e-9), k
    # and against the oracle on bf16-rounded logits
    q = out["query_embed"].detach().float().cpu().numpy()
    ps = out["candidate_embed"].pos.float().cpu().numpy()
    ng = out["candidate_embed"].neg.float().cpu().numpy()
    for name in orc.LOSS_NAMES:
        lb = None if name in orc.COSINE_LOSSES else "bf16"
        ref, _, _, _ = orc.lean_loss(name, q, ps, ng, orc.Config(**cfg_kw), with_grad=True, logits_dtype=lb)
        assert float(got[f"loss/{name}"]) == pytest.approx(ref, rel=4e-3, abs=4e-3), name


@pytest.mark.parametrize("tok_dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("graph", [True, False])
def test_pipelined_step_with_in_place_host_tokens_is_bitwise_identical(xr, tok_dtype, graph):
    """pipelined=True: load() runs the ingest phase (compaction, plan, gathers) on the copy stream and
    the gather reads pinned HOST token embeddings in place (only the selected rows cross PCIe);
    run() is the compute phase.  Same bits as the one-call step, for host and device inputs, across
    alternating step objects and repeated batches."""
    loss_fn = xr.InfoNCELoss(xr.LossConfig())
    batches = [orc.synth_batch(3000, 16, 60, dim=384, seed=s) for s in (5, 6, 7)]
    emb = xr.models.ItemEmbeddings(torch.from_numpy(batches[0]["table"]), add_padding_row=False).cuda()
    for b in batches[1:]:
        b["table"] = batches[0]["table"]
    plain = xr.PoolLossStep(emb, loss_fn, 16, 60, token_dtype=tok_dtype, use_graph=graph)
    pipes = [xr.PoolLossStep(emb, loss_fn, 16, 60, token_dtype=tok_dtype, use_graph=graph, pipelined=True,
                             host_tokens_in_place=bool(k)) for k in range(2)]   # one copies, one reads in place
    want = []
    for b in batches:
        l, g = run_step(plain, b, tok_dtype)
        want.append((l.clone(), g.clone()))
    host = [[torch.from_numpy(b["token_embeddings"]).to(tok_dtype).pin_memory()] +
            [torch.from_numpy(b[k]).pin_memory() for k in ("history_item_idx", "pos_item_idx", "neg_item_idx")]
            for b in batches]
    # software pipeline: the load (ingest) of batch i+1 is enqueued before the result of batch i is read
    pipes[0].load(*host[0])
    got = []
    for i in range(len(batches)):
        l, g = pipes[i % 2].run()
        if i + 1 < len(batches):
            pipes[(i + 1) % 2].load(*host[i + 1])
        got.append((l.clone(), g.clone()))
    torch.cuda.synchronize()
    for (l, g), (wl, wg) in zip(got, want):
        assert torch.equal(l, wl) and torch.equal(g, wg)
    # device-resident inputs through the pipelined object, and run() without a new load (full step)
    dev_args = [t.cuda() for t in host[1]]
    l, g = pipes[0](*dev_args)
    assert torch.equal(l, want[1][0]) and torch.equal(g, want[1][1])
    l2, g2 = pipes[0].run()
    assert torch.equal(l2, want[1][0]) and torch.equal(g2, want[1][1])


def test_step_orders_after_the_producer_stream(xr):
    """Device-resident inputs that are still being WRITTEN on the caller's stream when step(...) is
    called (the encoder output of the same training step): the copy stream must wait for the
    producer.  A long-running kernel delays the write of the token embeddings; without the
    ordering the step would read the stale (zero) buffer."""
    b = orc.synth_batch(2500, 12, 80, dim=384, seed=11)
    emb = xr.models.ItemEmbeddings(torch.from_numpy(b["table"]), add_padding_row=False).cuda()
    loss_fn = xr.InfoNCELoss(xr.LossConfig())
    want_loss, want_grad, _, _ = modular(xr, emb, loss_fn, b, torch.bfloat16)
    step = xr.PoolLossStep(emb, loss_fn, 12, 80)
    hist, pos, neg = (torch.from_numpy(b[k]).cuda() for k in
                      ("history_item_idx", "pos_item_idx", "neg_item_idx"))
    src = torch.from_numpy(b["token_embeddings"]).cuda().bfloat16()
    burn = torch.randn(4096, 4096, device="cuda")
    torch.cuda.synchronize()
    for _ in range(3):
        tok = torch.zeros_like(src)
        for _ in range(40):            # ~tens of ms of queued work ahead of the write below
            burn = (burn @ burn).clamp_(-1, 1)
        tok.copy_(src)                 # "the encoder" finishes on the current stream
        loss, dtok = step(tok, hist, pos, neg)
        del tok                        # the allocator may recycle the block: record_stream guards it
        torch.cuda.synchronize()
        assert torch.equal(loss, want_loss), (float(loss), float(want_loss))
        assert torch.equal(dtok.reshape(want_grad.shape), want_grad)


@pytest.mark.parametrize("name", ["NCELoss", "PairwiseHingeLoss", "PairwiseLogisticLoss"])
@pytest.mark.parametrize("cfg_kw", [{}, dict(mask_false_negatives=False, scale=3.0, margin=0.25)])
def test_one_pass_monitoring_serves_every_dot_family_train_loss(xr, name, cfg_kw):
    """The one-pass monitoring kernel with NCE / pairwise hinge / pairwise logistic (BPR) as the TRAIN loss:
    train loss and gradient bit-identical to the unmonitored step, every logged loss and statistic equal to the
    three-pass monitoring sequence (dot family and statistics to fp32 summation order, cosine family to bf16
    tolerance) and to the float64 oracle."""
    b = orc.synth_batch(3000, 16, 60, dim=384, seed=8)
    emb = xr.models.ItemEmbeddings(torch.from_numpy(b["table"]), add_padding_row=False).cuda()
    cfg = xr.LossConfig(**cfg_kw)
    loss_fn = getattr(xr, name)(cfg)
    plain = xr.PoolLossStep(emb, loss_fn, 16, 60)
    one = xr.PoolLossStep(emb, loss_fn, 16, 60, monitor=True)
    three = xr.PoolLossStep(emb, loss_fn, 16, 60, monitor=True, monitor_one_pass=False)
    assert one.monitor_one_pass and not three.monitor_one_pass
    l0, g0 = run_step(plain, b, torch.bfloat16)
    l1, g1 = run_step(one, b, torch.bfloat16)
    l3, g3 = run_step(three, b, torch.bfloat16)
    assert torch.equal(l0, l1) and torch.equal(g0, g1) and torch.equal(l0, l3) and torch.equal(g0, g3)
    got1, st1 = one.loss_dict()
    got3, st3 = three.loss_dict()
    for k, v in got3.items():
        tol = 4e-3 if k.split("/")[1] in orc.COSINE_LOSSES else 1e-6
        assert float(got1[k]) == pytest.approx(float(v), rel=tol, abs=tol), k
    for k, v in st3.items():
        assert st1[k] == pytest.approx(v, rel=1e-6, abs=1e-9), k
    assert float(got1[f"loss/{name}"]) == pytest.approx(float(l1), rel=1e-6)
