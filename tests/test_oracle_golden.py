"""Pin the numpy oracle against vectors produced by the reference's own losses.py
(tests/golden/make_golden.py).  CPU only."""

import json
import math

import numpy as np
import pytest

from oracle import xfmr_oracle as orc

CASES = [
    "anchor_m64_d384", "pool_default", "pool_nomask", "pool_scale20_margin02", "pool_margin0",
    "pool_hard5", "pool_hard5_nomask", "pool_autocast_bf16", "dense_first", "dense_diagonal",
    "dense_explicit", "dense_c1",
]


def load(golden_dir, name):
    z = np.load(golden_dir / f"losses_{name}.npz")
    cfg = orc.Config(**json.loads(str(z["cfg"])))
    return z, cfg


def dense_cand(z):
    if "cand" in z:
        return z["cand"]
    pos, neg = z["pos"], z["neg"]
    return np.concatenate([pos[:, None], np.broadcast_to(neg[None], (pos.shape[0],) + neg.shape)], 1)


@pytest.mark.parametrize("case", CASES)
def test_loss_from_reference_logits(golden_dir, case):
    """Given the reference's own logits the oracle's loss bodies must agree to fp32 rounding."""
    z, cfg = load(golden_dir, case)
    target = z["target"] if "target" in z else None
    for name in orc.LOSS_NAMES:
        logits = z["logits_cos"] if name in orc.COSINE_LOSSES else z["logits_dot"]
        tgt = orc.check_target(logits.shape[0], cfg, target)
        mask = orc.mine_hard_negatives(logits, orc.mask_false_negatives(logits, tgt, cfg), cfg)
        got = orc.loss_from_logits(name, logits, tgt, mask, cfg)
        want = float(z[f"loss/{name}"])
        # under autocast the reference evaluates parts of the loss body in bf16 as well
        rel = 2e-3 if case == "pool_autocast_bf16" else 2e-6
        assert got == pytest.approx(want, rel=rel, abs=2e-6), (case, name)


@pytest.mark.parametrize("case", CASES)
def test_stats_from_reference_logits(golden_dir, case):
    z, cfg = load(golden_dir, case)
    target = z["target"] if "target" in z else None
    logits = z["logits_dot"]
    tgt = orc.check_target(logits.shape[0], cfg, target)
    mask = orc.mine_hard_negatives(logits, orc.mask_false_negatives(logits, tgt, cfg), cfg)
    got = orc.logits_statistics(logits, tgt, mask, cfg)
    want = json.loads(str(z["stats"]))
    assert set(got) == set(want)
    for k, v in want.items():
        if math.isnan(v):
            assert math.isnan(got[k]), k
        else:
            # autocast: the reference reduces bf16 tensors, so its means are bf16-rounded
            tol = dict(rel=1e-2, abs=1e-3) if case == "pool_autocast_bf16" else dict(rel=1e-5, abs=1e-6)
            assert got[k] == pytest.approx(v, **tol), (case, k)


@pytest.mark.parametrize("case", [c for c in CASES if c != "pool_autocast_bf16"])
def test_end_to_end_fp32(golden_dir, case):
    """Oracle logits + loss + dL/dquery vs the reference (fp32): 1e-5 relative."""
    z, cfg = load(golden_dir, case)
    target = z["target"] if "target" in z else None
    cand = dense_cand(z)
    for name in orc.LOSS_NAMES:
        loss, dq = orc.embed_loss(name, z["query"], cand, cfg, target, with_grad=True)
        want = float(z[f"loss/{name}"])
        assert loss == pytest.approx(want, rel=1e-5, abs=1e-5), (case, name)
        ref_dq = z[f"dq/{name}"]
        scale = max(np.abs(ref_dq).max(), 1e-6)
        assert np.abs(dq - ref_dq).max() <= 2e-5 * scale + 1e-7, (case, name)


def test_autocast_bf16_dot_losses(golden_dir):
    """bf16-mixed: dot logits rounded to bf16 before masking."""
    z, cfg = load(golden_dir, "pool_autocast_bf16")
    cand = dense_cand(z)
    ref_logits = z["logits_dot"]
    q = orc.round_bf16(z["query"]).astype(np.float64)
    c = orc.round_bf16(cand).astype(np.float64)
    mine = orc.round_bf16(orc.dot_logits(q, c).astype(np.float32))
    # accumulation order differs from MKL's bf16 gemm: allow one bf16 ulp on a few entries
    assert np.mean(mine != ref_logits) < 0.02
    assert np.abs(mine - ref_logits).max() <= 2 ** -7 * np.abs(ref_logits).max()
    for name in ["InfoNCELoss", "NCELoss", "PairwiseHingeLoss", "PairwiseLogisticLoss"]:
        tgt = orc.check_target(ref_logits.shape[0], cfg, None)
        mask = orc.mask_false_negatives(ref_logits, tgt, cfg)
        got = orc.loss_from_logits(name, ref_logits, tgt, mask, cfg)
        assert got == pytest.approx(float(z[f"loss/{name}"]), rel=2e-3), name


@pytest.mark.parametrize("case", ["anchor_m64_d384", "pool_default", "pool_scale20_margin02", "pool_nomask"])
def test_lean_equals_dense(golden_dir, case):
    """[rowdot | Q.Neg^T] form == dense (M,1+M,D) form."""
    z, cfg = load(golden_dir, case)
    for name in orc.LOSS_NAMES:
        loss, dq, _, _ = orc.lean_loss(name, z["query"], z["pos"], z["neg"], cfg, with_grad=True)
        assert loss == pytest.approx(float(z[f"loss/{name}"]), rel=1e-5, abs=1e-5)
        ref_dq = z[f"dq/{name}"]
        assert np.abs(dq - ref_dq).max() <= 2e-5 * max(np.abs(ref_dq).max(), 1e-6) + 1e-7


def test_survey_anchor_values(golden_dir):
    z, _ = load(golden_dir, "anchor_m64_d384")
    want = {"InfoNCELoss": 36.929882, "PairwiseLogisticLoss": 17.622574, "PairwiseHingeLoss": 16.084221,
            "NCELoss": 624.552917, "AlignmentLoss": 64.235138, "AlignmentContrastiveLoss": 64.235138,
            "ContrastiveLoss": 0.0}
    for k, v in want.items():
        assert float(z[f"loss/{k}"]) == pytest.approx(v, rel=1e-6, abs=1e-6)


def test_round_bf16_matches_torch():
    import torch

    x = np.random.default_rng(0).standard_normal(4096).astype(np.float32) * 3
    assert np.array_equal(orc.round_bf16(x), torch.from_numpy(x).bfloat16().float().numpy())


def test_compute_embeds_matches_torch_restatement():
    """compute_embeds restatement vs the same lines written with torch ops (models.py:388-419)."""
    import torch

    b = orc.synth_batch(50, 4, 9, dim=16, seed=3)
    out = orc.compute_embeds(b["table"], b["token_embeddings"], b["history_item_idx"],
                             b["pos_item_idx"], b["neg_item_idx"])
    emb = torch.nn.Embedding.from_pretrained(torch.from_numpy(b["table"]), freeze=True, padding_idx=0)
    hist, pos, neg = (torch.from_numpy(b[k]) for k in ("history_item_idx", "pos_item_idx", "neg_item_idx"))
    am = (emb(hist) != 0).any(-1)
    q = torch.from_numpy(b["token_embeddings"])[am]
    p = emb(pos[am])[:, None, :]
    n = emb(neg[am])[None].expand(p.size(0), -1, -1)
    cand = torch.cat([p, n], 1)
    pm = pos[am] != 0
    assert np.array_equal(out["query_embed"], q[pm].numpy())
    assert np.array_equal(out["candidate_embed"], cand[pm].numpy())
    assert np.array_equal(out["positive_mask"], pm.numpy())
    assert out["query_embed"].shape[0] > 0 and (~pm).any()


def test_exact_search_tie_rule_and_exclusion():
    cat = np.zeros((10, 4), np.float32)
    cat[:, 0] = 1.0          # all identical -> all scores tie
    cat[7] = [0, 1, 0, 0]
    s, i = orc.exact_search(np.array([1, 0, 0, 0], np.float32), cat, 4, exclude=[[0, 2]])
    assert i.tolist() == [[1, 3, 4, 5]]
    s, i = orc.exact_search(np.array([1, 0, 0, 0], np.float32), cat, 3, exclude=None, chunk=4)
    assert i.tolist() == [[0, 1, 2]]


def test_topk_rows_stable():
    s = np.array([[1, 3, 3, 2, 3]], np.float32)
    _, idx = orc.topk_rows(s, 2)
    assert idx.tolist() == [[1, 2]]


def test_retrieval_metrics_known_answers():
    assert orc.retrieval_metrics(["a"], [], 3) == {}
    m = orc.retrieval_metrics(["i10", "i3", "i7"], {"i3"}, 3)   # docstring example metrics.py:56-60
    assert m["retrieval_auroc"] == pytest.approx(0.5)
    assert m["retrieval_reciprocal_rank"] == pytest.approx(0.5)
    assert m["retrieval_precision"] == pytest.approx(1 / 3)
    assert m["retrieval_recall"] == 1.0 and m["retrieval_hit_rate"] == 1.0
    assert m["retrieval_normalized_dcg"] == pytest.approx(1 / math.log2(3))
    assert m["retrieval_average_precision"] == pytest.approx(0.5)
    # target missing from the list: appended after the padded recs, outside top_k
    m = orc.retrieval_metrics(["a", "b"], {"z"}, 4)
    assert all(v == 0.0 for v in m.values())
    # two hits among 4, 3 targets in total
    m = orc.retrieval_metrics(["a", "b", "c", "d"], {"a", "c", "q"}, 4)
    assert m["retrieval_recall"] == pytest.approx(2 / 3)
    assert m["retrieval_precision"] == pytest.approx(0.5)
    assert m["retrieval_average_precision"] == pytest.approx((1 / 1 + 2 / 3) / 2)
    assert m["retrieval_auroc"] == pytest.approx(3 / 4)
    idcg = 1 + 1 / math.log2(3) + 1 / math.log2(4)
    assert m["retrieval_normalized_dcg"] == pytest.approx((1 + 1 / math.log2(4)) / idcg)


EMBED_CASES = ["basic", "normalized", "truncated"]


@pytest.mark.parametrize("case", EMBED_CASES)
def test_compute_embeds_oracle_matches_reference_outputs(golden_dir, case):
    """The oracle's compute_embeds against the outputs of the reference's OWN
    RecommenderModel.forward + compute_embeds (models.py:306-345, 366-419), executed from the
    reference's source file by tests/golden/make_golden_embeds.py."""
    z = np.load(golden_dir / f"embeds_{case}.npz")
    out = orc.compute_embeds(z["table"], z["tokens"], z["history_item_idx"], z["pos_item_idx"],
                             z["neg_item_idx"], dense=True, is_normalized=bool(z["is_normalized"]))
    assert np.array_equal(out["attention_mask"], z["attention_mask"].astype(bool))
    assert np.array_equal(out["positive_mask"], z["positive_mask"].astype(bool))
    assert np.array_equal(out["candidate_embed"], z["candidate_embed"])          # gathered rows: bit-exact
    if bool(z["is_normalized"]):
        np.testing.assert_allclose(out["query_embed"], z["query_embed"], rtol=1e-6, atol=1e-7)
    else:
        assert np.array_equal(out["query_embed"], z["query_embed"])
