"""SeqBatch construction (SeqDataset.__getitem__ + collate, data.py:669-805).

CPU: the oracle's Philox against the Random123 known-answer vectors, the oracle's examples against
the reference's support constraints, distribution checks.  GPU: the kernel bit-exact against the
oracle, the constraints at scale, uniformity, batch-composition independence."""

import numpy as np
import pytest
import torch

from oracle import xfmr_oracle as orc


def make_histories(rng, n_users, n_items, max_len, p_pos=0.6, dup=True):
    hs, ls = [], []
    for _ in range(n_users):
        n = int(rng.integers(0, max_len + 1))
        h = rng.integers(1, n_items + 1, size=n)
        if not dup and n <= n_items:
            h = rng.permutation(n_items)[:n] + 1
        l = rng.random(n) < p_pos
        if n:   # map_id2idx trims events after the last positive (data.py:609-612)
            last = np.flatnonzero(l).max(initial=-1) + 1
            h, l = h[:last], l[:last]
        hs.append(h.astype(np.int64))
        ls.append(l)
    return hs, ls


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert orc.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert orc.philox4x32_10((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert orc.philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == (
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


@pytest.mark.parametrize("n_items,max_len,L,look", [(50, 30, 8, 0), (50, 30, 8, 3), (7, 12, 5, 0),
                                                    (1000, 120, 32, 0), (3, 6, 4, 1)])
def test_oracle_examples_satisfy_reference_constraints(n_items, max_len, L, look):
    rng = np.random.default_rng(n_items + L)
    hs, ls = make_histories(rng, 40, n_items, max_len)
    for step in range(3):
        for u, (h, l) in enumerate(zip(hs, ls)):
            if len(h) == 0:
                continue
            ho, po, no = orc.seq_sample_example(h, l, u, n_items, L, look, seed=5, step=step)
            orc.check_seq_example(h, l, ho, po, no, n_items, L, look)


def _chi2(counts, expected):
    return float(((counts - expected) ** 2 / expected).sum())


def test_oracle_distributions_are_uniform():
    """positions: every position is kept with probability L/cand; negatives: uniform over the
    complement of the history; positives: uniform over the window's positives."""
    n_items, L = 40, 6
    h = np.array([3, 9, 9, 14, 21, 3, 30, 31, 32, 33, 35, 36, 38, 2, 1], np.int64)   # cand = 14 > L
    l = np.ones(len(h), bool)
    l[[1, 4]] = False
    trials = 3000
    pos_cnt = np.zeros(len(h) - 1)
    neg_cnt = np.zeros(n_items + 1)
    first_pos = {}
    for step in range(trials):
        ho, po, no = orc.seq_sample_example(h, l, 7, n_items, L, 0, seed=1, step=step)
        # distinct-item positions are identifiable; count by matching value+order greedily
        j = 0
        for v in ho:
            while h[j] != v:
                j += 1
            pos_cnt[j] += 1
            j += 1
        neg_cnt[no] += 1
        first_pos[int(po[-1])] = first_pos.get(int(po[-1]), 0) + 1
    comp = np.setdiff1d(np.arange(1, n_items + 1), h)
    e = trials * L / len(comp)
    assert _chi2(neg_cnt[comp], e) < 2.0 * len(comp)          # dof ~ 27; generous bound
    assert neg_cnt[np.unique(h)].sum() == 0 and neg_cnt[0] == 0
    # duplicates (items 3 and 9 twice) blur per-position counts between the twins; check totals
    assert abs(pos_cnt.sum() - trials * L) < 1e-9
    singles = [i for i in range(len(h) - 1) if (h[:-1] == h[i]).sum() == 1]
    e = trials * L / (len(h) - 1)
    assert _chi2(pos_cnt[singles], e) < 3.0 * len(singles)


# ------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def xr():
    import xfmr_rec_b200 as pkg

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return pkg


def _sampler(xr, hs, ls, n_items, L, look, seed=3):
    return xr.data.SeqBatchSampler(xr.data.SeqDataConfig(L, look), hs, ls, n_items, device="cuda", seed=seed)


@pytest.mark.gpu
@pytest.mark.parametrize("n_items,max_len,L,look", [(50, 30, 8, 0), (50, 30, 8, 3), (7, 12, 5, 0), (3, 6, 4, 1),
                                                    (2000, 700, 32, 0), (300, 900, 200, 5), (100000, 400, 50, 0)])
def test_kernel_is_bit_exact_against_the_oracle(xr, n_items, max_len, L, look):
    rng = np.random.default_rng(n_items * 7 + L)
    hs, ls = make_histories(rng, 60, n_items, max_len)
    s = _sampler(xr, hs, ls, n_items, L, look)
    kept_h = [hs[i] for i in s.kept_histories]
    kept_l = [ls[i] for i in s.kept_histories]
    rows = torch.from_numpy(rng.permutation(len(s))).cuda()
    for step in (0, 1, 12345678901):
        got = s.sample(rows, step=step, return_lengths=True)
        want = orc.seq_sample_batch(kept_h, kept_l, rows.cpu().numpy(), n_items, L, look, seed=3, step=step,
                                    row_hist=s.row_hist.cpu().numpy())
        for k in ("history_item_idx", "pos_item_idx", "neg_item_idx", "seq_len"):
            assert np.array_equal(got[k].cpu().numpy(), want[k]), (k, step)


@pytest.mark.gpu
def test_kernel_respects_reference_constraints_at_scale(xr):
    rng = np.random.default_rng(0)
    n_items, L = 3706, 50          # MovieLens-1M-shaped (BASELINE configs[0])
    hs, ls = make_histories(rng, 600, n_items, 400)
    s = _sampler(xr, hs, ls, n_items, L, 0)
    rh = s.row_hist.cpu().numpy()
    rows = torch.arange(len(s), device="cuda")
    got = {k: v.cpu().numpy() for k, v in s.sample(rows, step=9, return_lengths=True).items()}
    for b in range(len(s)):
        h = s.kept_histories[rh[b]]
        n = got["seq_len"][b]
        assert (got["history_item_idx"][b, n:] == 0).all() and (got["neg_item_idx"][b, n:] == 0).all()
        orc.check_seq_example(hs[h], ls[h], got["history_item_idx"][b, :n], got["pos_item_idx"][b, :n],
                              got["neg_item_idx"][b, :n], n_items, L, 0)


@pytest.mark.gpu
def test_kernel_rows_do_not_depend_on_batch_composition(xr):
    rng = np.random.default_rng(2)
    hs, ls = make_histories(rng, 80, 500, 120)
    s = _sampler(xr, hs, ls, 500, 16, 0)
    rows = torch.arange(len(s), device="cuda")
    full = s.sample(rows, step=4)
    sub = s.sample(rows[5:9], step=4)
    for k in full:
        assert torch.equal(full[k][5:9], sub[k])
    again = s.sample(rows, step=4)
    other = s.sample(rows, step=5)
    assert all(torch.equal(full[k], again[k]) for k in full)
    assert not torch.equal(full["neg_item_idx"], other["neg_item_idx"])


@pytest.mark.gpu
def test_kernel_negatives_are_uniform(xr):
    n_items, L = 64, 20
    h = np.arange(1, 41, dtype=np.int64)              # cand = 39 > L, complement = items 41..64 (24 >= L)
    l = np.ones(40, bool)
    s = _sampler(xr, [h], [l], n_items, L, 0)
    rows = torch.zeros(1, dtype=torch.int64, device="cuda")
    neg_cnt = np.zeros(n_items + 1)
    pos_cnt = np.zeros(n_items + 1)
    trials = 4000
    for step in range(trials):
        out = s.sample(rows, step=step)
        neg_cnt[out["neg_item_idx"][0].cpu().numpy()] += 1
        pos_cnt[out["history_item_idx"][0].cpu().numpy()] += 1
    assert neg_cnt[:41].sum() == 0
    e = trials * L / 24
    assert _chi2(neg_cnt[41:], e) < 60          # dof 23 (p ~ 4e-5 at 60 for independent draws)
    e = trials * L / 39
    assert pos_cnt[40] == 0 and _chi2(pos_cnt[1:40], e) < 90


@pytest.mark.gpu
def test_sampled_batch_feeds_the_train_step(xr):
    """The device-built SeqBatch goes straight into the scoring-and-loss step (no H2D of indices)."""
    if torch.cuda.get_device_capability()[0] != 10:
        pytest.skip("needs an sm_100 device")
    rng = np.random.default_rng(4)
    n_items, L, B = 3000, 32, 16
    hs, ls = make_histories(rng, 200, n_items, 150)
    s = _sampler(xr, hs, ls, n_items, L, 0)
    batch = next(iter(s.epoch(B, epoch=0)))
    table = torch.randn((n_items + 1, 384), device="cuda") / 384 ** 0.5
    table[0] = 0
    emb = xr.models.ItemEmbeddings(table, add_padding_row=False).cuda()
    tok = (torch.randn((B, L, 384), device="cuda") / 384 ** 0.5).bfloat16().requires_grad_(True)
    out = xr.models.compute_embeds(emb, tok, batch["history_item_idx"], batch["pos_item_idx"],
                                   batch["neg_item_idx"], candidate_dtype=torch.bfloat16)
    loss = xr.InfoNCELoss(xr.LossConfig())(out["query_embed"], out["candidate_embed"])
    loss.backward()
    assert torch.isfinite(loss) and float(loss) > 0 and tok.grad is not None


# ---------------------------------------------------------------- against the reference's own sampler
def _two_sample_chi2(a, b):
    """Pearson two-sample statistic and its degrees of freedom over the cells either sample hit."""
    a, b = np.asarray(a, float).ravel(), np.asarray(b, float).ravel()
    keep = (a + b) > 0
    a, b = a[keep], b[keep]
    A, B = a.sum(), b.sum()
    if A == 0 or B == 0 or keep.sum() < 2:
        return 0.0, 0
    stat = (((np.sqrt(B / A) * a - np.sqrt(A / B) * b) ** 2) / (a + b)).sum()
    return float(stat), int(keep.sum() - 1)


def test_reference_examples_satisfy_the_support_constraints(golden_dir):
    """Everything the reference's OWN sample_sequence / sample_positives / sample_negatives produced
    (data.py:669-747, executed by tests/golden/make_golden_seqbatch.py) passes check_seq_example: the
    constraints the device sampler is tested against are not stricter than the reference."""
    z = np.load(golden_dir / "seqbatch_reference_stats.npz")
    n_items, L = int(z["n_items"]), int(z["max_seq_length"])
    for look in (0, 3):
        for u in range(6):
            h, l = z[f"hist{u}"], z[f"label{u}"]
            for ex in z[f"look{look}_user{u}_raw"]:
                n = int((ex[0] >= 0).sum())
                orc.check_seq_example(h, l, ex[0, :n], ex[1, :n], ex[2, :n], n_items, L, look)


@pytest.mark.parametrize("look", [0, 3])
def test_sampler_distribution_matches_the_reference(golden_dir, look):
    """Marginal distributions of the counter-based algorithm (what the kernel computes, bit for bit)
    against 6,000 draws of the reference's own sampler: which positions are kept, which positive each
    kept position gets, which negatives are drawn — including the with-replacement case (fewer
    candidates than positions) and a history that covers the whole catalog."""
    z = np.load(golden_dir / "seqbatch_reference_stats.npz")
    n_items, L, draws = int(z["n_items"]), int(z["max_seq_length"]), 1500
    for u in range(6):
        h, l = z[f"hist{u}"], z[f"label{u}"]
        pos_cnt = np.zeros(len(h), np.int64)
        positive_cnt = np.zeros((len(h), n_items + 1), np.int64)
        neg_cnt = np.zeros(n_items + 1, np.int64)
        distinct = len(set(h.tolist())) == len(h)
        for step in range(draws):
            ho, po, no = orc.seq_sample_example(h, l, row=u, n_items=n_items, max_seq_length=L,
                                                pos_lookahead=look, seed=11, step=step)
            np.add.at(neg_cnt, no, 1)
            if distinct:     # positions are identifiable from the items
                idx = [int(np.flatnonzero(h == v)[0]) for v in ho]
                pos_cnt[idx] += 1
                positive_cnt[idx, po] += 1
        key = f"look{look}_user{u}"
        stat, dof = _two_sample_chi2(neg_cnt, z[f"{key}_negatives"])
        assert stat <= dof + 5 * np.sqrt(2 * max(dof, 1)) + 5, (key, "negatives", stat, dof)
        if distinct:
            stat, dof = _two_sample_chi2(pos_cnt, z[f"{key}_positions"])
            assert stat <= dof + 5 * np.sqrt(2 * max(dof, 1)) + 5, (key, "positions", stat, dof)
            # support of the positives is identical; frequencies agree
            assert ((positive_cnt > 0) <= (z[f"{key}_positives"] > 0)).all(), key
            stat, dof = _two_sample_chi2(positive_cnt, z[f"{key}_positives"])
            assert stat <= dof + 5 * np.sqrt(2 * max(dof, 1)) + 5, (key, "positives", stat, dof)


def test_history_csr_layout_follows_process_events():
    """build_history_csr = the host-side part of SeqDataset.process_events (data.py:638-657): empty
    histories dropped, one dataset row per started block of max_seq_length events (duplicate_rows,
    data.py:618-636), per-history prefix counts of positive labels and sorted unique items."""
    from xfmr_rec_b200.data import build_history_csr

    hs = [np.array([5, 3, 5, 9]), np.array([], np.int64), np.array([2]), np.arange(1, 12)]
    ls = [np.array([1, 0, 1, 1], bool), np.array([], bool), np.array([1], bool), np.ones(11, bool)]
    c = build_history_csr(hs, ls, num_items=20, max_seq_length=4)
    assert c["kept"] == [0, 2, 3]
    assert c["hist_off"].tolist() == [0, 4, 5, 16]
    assert c["items"].tolist() == [5, 3, 5, 9, 2] + list(range(1, 12))
    assert c["pos_prefix"].tolist()[:5] == [1, 1, 2, 3, 1]
    assert c["uniq_off"].tolist() == [0, 3, 4, 15] and c["uniq_items"].tolist()[:4] == [3, 5, 9, 2]
    # rows: ceil-style (len - 1) // L + 1  ->  1, 1, 3
    assert c["row_hist"].tolist() == [0, 1, 2, 2, 2]
    with pytest.raises(IndexError):
        build_history_csr([np.array([0, 1])], [np.array([1, 1], bool)], 20, 4)
    with pytest.raises(IndexError):
        build_history_csr([np.array([21])], [np.array([1], bool)], 20, 4)
