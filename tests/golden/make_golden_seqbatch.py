"""Golden statistics for SeqBatch construction, produced by EXECUTING the reference's own
``SeqDataset.sample_sequence / sample_positives / sample_negatives`` (xfmr_rec/data.py:669-747) —
build container only:

    python tests/golden/make_golden_seqbatch.py

``xfmr_rec/data.py`` cannot be imported here (polars / lightning / bentoml absent); the three method
definitions are pure numpy, so they are taken from the source file where it lies (``ast``; nothing is
copied into this repo) and bound to a stand-in object with the attributes they touch (``rng``,
``config.max_seq_length``, ``config.pos_lookahead``, ``all_idx``).  The reference draws from an
unseeded generator (data.py:574), so no value-level golden exists; what is stored is (a) raw examples
(to check that the support constraints of oracle.check_seq_example accept everything the reference
produces) and (b) marginal histograms over many draws (to compare distributions).
"""

from __future__ import annotations

import ast
import pathlib
import types

import numpy as np

SRC = pathlib.Path("/root/reference/xfmr_rec/data.py")
OUT = pathlib.Path(__file__).parent
N_ITEMS, L, DRAWS, RAW = 40, 6, 6000, 40


def reference_methods():
    tree = ast.parse(SRC.read_text())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "SeqDataset")
    names = ("sample_sequence", "sample_positives", "sample_negatives")
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in names]
    for f in fns:            # annotations reference typing aliases of the module: not needed to run
        f.returns = None
        for a in f.args.args + f.args.kwonlyargs:
            a.annotation = None
    ns = {"np": np}
    exec(compile(ast.Module(body=fns, type_ignores=[]), str(SRC), "exec"), ns)
    return [ns[n] for n in names]


def histories():
    rng = np.random.default_rng(0)
    hs = [
        np.array([3, 9, 14, 21, 30, 31, 32, 33, 35, 36, 38, 2, 1, 17]),      # long, distinct
        np.array([5, 5, 7, 9, 5, 11, 7, 13, 15, 5, 19]),                      # long, duplicates
        np.array([4, 8, 12, 16]),                                             # short: taken whole
        np.array([6, 10]),                                                    # one position
        rng.permutation(N_ITEMS)[:37] + 1,                                    # 3 negative candidates < seq_len
        np.arange(1, N_ITEMS + 1),                                            # history covers the catalog
    ]
    ls = [rng.random(len(h)) < 0.6 for h in hs]
    for l in ls:
        l[-1] = True      # map_id2idx trims after the last positive (data.py:609-612)
    return [h.astype(np.int64) for h in hs], ls


if __name__ == "__main__":
    sample_sequence, sample_positives, sample_negatives = reference_methods()
    hs, ls = histories()
    rec = {"n_items": np.array(N_ITEMS), "max_seq_length": np.array(L), "draws": np.array(DRAWS)}
    for u, (h, l) in enumerate(zip(hs, ls)):
        rec[f"hist{u}"], rec[f"label{u}"] = h, l
    for look in (0, 3):
        me = types.SimpleNamespace(rng=np.random.default_rng(100 + look), all_idx=set(range(1, N_ITEMS + 1)),
                                   config=types.SimpleNamespace(max_seq_length=L, pos_lookahead=look))
        for u, (h, l) in enumerate(zip(hs, ls)):
            pos_cnt = np.zeros(len(h), np.int64)              # how often each POSITION was sampled
            positive_cnt = np.zeros((len(h), N_ITEMS + 1), np.int64)   # per sampled position: chosen positive
            neg_cnt = np.zeros(N_ITEMS + 1, np.int64)
            raw = []
            for d in range(DRAWS):
                idx = sample_sequence(me, h)
                p = sample_positives(me, history_item_idx=h, history_label=l, sampled_indices=idx)
                n = sample_negatives(me, history_item_idx=h, sampled_indices=idx)
                pos_cnt[idx] += 1
                positive_cnt[idx, p] += 1
                np.add.at(neg_cnt, n, 1)
                if d < RAW:
                    raw.append(np.stack([h[idx], p, n]) if len(idx) else np.zeros((3, 0), np.int64))
            key = f"look{look}_user{u}"
            rec[f"{key}_positions"] = pos_cnt
            rec[f"{key}_positives"] = positive_cnt
            rec[f"{key}_negatives"] = neg_cnt
            width = max(r.shape[1] for r in raw)
            rec[f"{key}_raw"] = np.stack([np.pad(r, ((0, 0), (0, width - r.shape[1])), constant_values=-1)
                                          for r in raw])
            print(key, "positions", pos_cnt.tolist())
    np.savez_compressed(OUT / "seqbatch_reference_stats.npz", **rec)
