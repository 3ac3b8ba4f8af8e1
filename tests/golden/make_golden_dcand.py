"""Golden vectors for dL/d candidate_embed, produced by EXECUTING the reference's own loss module.

Run in the build container only (``/root/reference`` is not on the GPU box):

    python tests/golden/make_golden_dcand.py

The reference's losses are differentiable in both arguments (xfmr_rec/losses.py:128-155).  For a dense
(M, C, D) candidate tensor with ``requires_grad`` this stores, per loss class, the loss value, dL/dquery and
dL/dcandidate_embed for two configurations (default; scale / margin / no false-negative mask, explicit
targets).  Only the reference's outputs are stored.
"""
from __future__ import annotations

import json
import pathlib
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from xfmr_rec import losses as ref  # noqa: E402

OUT = pathlib.Path(__file__).parent


def main():
    torch.manual_seed(7)
    m, c, d = 9, 6, 24
    q = torch.randn(m, d) * d ** -0.5
    cand = torch.randn(m, c, d) * d ** -0.5
    cand[2, 3] = cand[2, 0]          # a negative identical to the row's positive (tie => masked)
    cand[4, 1] = 0.0                 # zero-norm candidate (cosine clamp)
    cases = {
        "default": ({}, None),
        "explicit_scaled": ({"target_position": None, "scale": 3.0, "margin": 0.3, "mask_false_negatives": False},
                            torch.tensor([0, 5, 2, 1, 3, 4, 0, 2, 5])),
    }
    for tag, (kw, target) in cases.items():
        cfg = ref.LossConfig(**kw)
        rec = {"query": q.numpy(), "cand": cand.numpy(), "cfg": json.dumps(kw)}
        if target is not None:
            rec["target"] = target.numpy()
        for cls in ref.LOSS_CLASSES:
            qq, cc = q.clone().requires_grad_(True), cand.clone().requires_grad_(True)
            loss = cls(cfg)(query_embed=qq, candidate_embed=cc, target=target)
            loss.backward()
            rec[f"loss/{cls.__name__}"] = np.array(loss.item(), dtype=np.float64)
            rec[f"dq/{cls.__name__}"] = qq.grad.numpy()
            rec[f"dcand/{cls.__name__}"] = cc.grad.numpy()
        np.savez_compressed(OUT / f"losses_dcand_{tag}.npz", **rec)
        print(tag, {k: float(v) for k, v in rec.items() if k.startswith("loss/")})


if __name__ == "__main__":
    main()
