"""Run HERE (build container, /root/reference mounted): the reference's OWN code path against the
lean port that bench.py times on the GPU box, on the same CPU, same inputs.

The verbatim path = models.py:398-416 candidate materialisation (expand + cat -> (M, 1+M_a, D)) +
xfmr_rec.losses.InfoNCELoss (imported from /root/reference, unmodified), forward + backward to the
query rows.  The lean port = oracle/cpu_baseline.train_step.  Writes profiles/cpu_reference_r01.json
(a committed fixture: /root/reference does not exist on the GPU box).

    PYTHONPATH=/root/reference python oracle/time_reference_verbatim.py

TEST / MEASUREMENT INFRASTRUCTURE ONLY."""
import json
import os
import pathlib
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), "/root/reference"]
import torch

from oracle import cpu_baseline, xfmr_oracle as orc
from xfmr_rec import losses as ref_losses   # the reference, unmodified

torch.set_num_threads(os.cpu_count() or 1)
out = []
for batch_size, seq_len, n_items in ((16, 50, 3706), (32, 50, 3706), (8, 200, 27278)):
    b = orc.synth_batch(n_items, batch_size, seq_len, dim=384, seed=0)
    table, tokens, hist, pos, neg = (torch.from_numpy(b[k]) for k in
                                     ("table", "token_embeddings", "history_item_idx", "pos_item_idx",
                                      "neg_item_idx"))

    def verbatim():
        tok = tokens.detach().requires_grad_(True)
        input_embeds = table[hist]                                   # models.py:336-338
        attention_mask = (input_embeds != 0).any(-1)                 # models.py:343
        query = tok[attention_mask]                                  # models.py:392
        pos_sel = pos[attention_mask]
        pos_embed = table[pos_sel]                                   # models.py:400
        neg_embed = table[neg[attention_mask]]                       # models.py:406
        cand = torch.cat([pos_embed[:, None, :],                     # models.py:408-410
                          neg_embed[None, :, :].expand(query.size(0), -1, -1)], dim=1)
        pos_mask = pos_sel != 0                                      # models.py:413
        loss = ref_losses.InfoNCELoss(ref_losses.LossConfig())(     # losses.py:128-155, 479-488
            query_embed=query[pos_mask], candidate_embed=cand[pos_mask])
        loss.backward()
        return float(loss.detach()), tok.grad, int(pos_mask.sum()), cand.shape

    def lean():
        return cpu_baseline.train_step(table, tokens, hist, pos, neg)

    lv, gv, m, shape = verbatim()
    ll, gl = lean()
    rel = abs(lv - ll) / abs(lv)
    gerr = float((gv - gl).abs().max() / gv.abs().max())

    def best_of(fn, reps=3):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return min(ts)

    tv, tl = best_of(verbatim), best_of(lean)
    rec = {"batch": batch_size, "seq_len": seq_len, "items": n_items, "rows_M": m,
           "candidate_tensor": list(shape), "candidate_tensor_GB": shape[0] * shape[1] * shape[2] * 4 / 1e9,
           "cores": os.cpu_count(), "reference_verbatim_ms": tv * 1e3, "lean_port_ms": tl * 1e3,
           "verbatim_over_lean": tv / tl, "loss_rel_diff": rel, "grad_max_rel_diff": gerr}
    out.append(rec)
    print(json.dumps(rec))
(ROOT / "profiles" / "cpu_reference_r01.json").write_text(json.dumps(out, indent=1))
