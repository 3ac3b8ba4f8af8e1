"""Generate golden vectors by EXECUTING the reference's own loss module.

Run in the build container only (``/root/reference`` is not on the GPU box):

    python tests/golden/make_golden.py

Imports ``xfmr_rec/losses.py`` from the read-only reference checkout (its only
imports are pydantic + torch), feeds it seeded inputs built the way
``xfmr_rec/models.py:398-416`` builds them, and stores inputs + the reference's
outputs (loss values, autograd dL/dquery, LogitsStatistics) as small ``.npz``
fixtures.  Nothing from the reference is copied; only its outputs are stored.
"""

from __future__ import annotations

import json
import pathlib
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
from xfmr_rec import losses as ref  # noqa: E402

OUT = pathlib.Path(__file__).parent
LOSS_NAMES = [cls.__name__ for cls in ref.LOSS_CLASSES]


def dense_candidates(pos, neg):
    # models.py:402-410
    return torch.cat([pos[:, None, :], neg[None, :, :].expand(pos.size(0), -1, -1)], dim=1)


def run_case(name, q, cand, cfg_kwargs, target=None, autocast=False, extra=None):
    cfg = ref.LossConfig(**cfg_kwargs)
    rec = {"query": q.numpy(), "cfg": json.dumps(cfg_kwargs), "autocast": np.array(autocast)}
    if extra:
        rec.update(extra)
    else:
        rec["cand"] = cand.numpy()
    if target is not None:
        rec["target"] = target.numpy()
    for loss_name, cls in zip(LOSS_NAMES, ref.LOSS_CLASSES):
        qq = q.clone().requires_grad_(True)
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
            loss = cls(cfg)(query_embed=qq, candidate_embed=cand, target=target)
        loss.backward()
        rec[f"loss/{loss_name}"] = np.array(loss.item(), dtype=np.float64)
        rec[f"dq/{loss_name}"] = qq.grad.numpy()
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        stats = ref.LogitsStatistics(cfg)(query_embed=q, candidate_embed=cand, target=target)
        dot = ref.InfoNCELoss(cfg).compute_logits(q, cand)
        cos = ref.InfoNCELoss(cfg).cosine_similarity_logits(q, cand)
    rec["stats"] = json.dumps(stats)
    rec["logits_dot"] = dot.float().numpy()
    rec["logits_cos"] = cos.float().numpy()
    np.savez_compressed(OUT / f"losses_{name}.npz", **rec)
    print(name, {k: float(v) for k, v in rec.items() if k.startswith("loss/")})


def main():
    # --- SURVEY §8(c) sanity anchor: M=64, D=384, default LossConfig ----------------
    torch.manual_seed(0)
    m, d = 64, 384
    q, pos, neg = torch.randn(m, d), torch.randn(m, d), torch.randn(m, d)
    run_case("anchor_m64_d384", q, dense_candidates(pos, neg), {},
             extra={"pos": pos.numpy(), "neg": neg.numpy()})

    # --- scaled (1/sqrt(D)) shared pool, several configs ---------------------------
    torch.manual_seed(1)
    m, d = 40, 96
    sc = d ** -0.5
    q, pos, neg = torch.randn(m, d) * sc, torch.randn(m, d) * sc, torch.randn(m + 7, d) * sc
    neg[3] = pos[5]          # duplicate of a positive inside the pool (tie => masked)
    neg[11] = 0.0            # padding row in the pool (zero norm)
    for tag, kw in {
        "default": {},
        "nomask": {"mask_false_negatives": False},
        "scale20_margin02": {"scale": 20.0, "margin": 0.2},
        "margin0": {"margin": 0.0},
        "hard5": {"num_hard_negatives": 5},
        "hard5_nomask": {"num_hard_negatives": 5, "mask_false_negatives": False},
    }.items():
        run_case(f"pool_{tag}", q, dense_candidates(pos, neg), kw,
                 extra={"pos": pos.numpy(), "neg": neg.numpy()})

    # --- bf16-mixed autocast (trainer.py:450) ----------------------------------------
    run_case("pool_autocast_bf16", q, dense_candidates(pos, neg), {}, autocast=True,
             extra={"pos": pos.numpy(), "neg": neg.numpy()})

    # --- genuine per-row dense candidates, explicit / diagonal targets ---------------
    torch.manual_seed(2)
    m, c, d = 24, 24, 64
    q, cand = torch.randn(m, d) * d ** -0.5, torch.randn(m, c, d) * d ** -0.5
    run_case("dense_first", q, cand, {})
    run_case("dense_diagonal", q, cand, {"target_position": "diagonal"})
    tgt = torch.randint(0, c, (m,))
    run_case("dense_explicit", q, cand, {"target_position": None}, target=tgt)
    # single-candidate rows: no negatives at all (std of one element = nan, empty neg block)
    run_case("dense_c1", q[:1], cand[:1, :1], {})


if __name__ == "__main__":
    main()
