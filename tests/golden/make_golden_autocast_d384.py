"""Golden vectors of the reference's own bf16-mixed autocast path at D = 384 (the dimension the tensor-core
kernels are specialised for), by EXECUTING xfmr_rec/losses.py under ``torch.autocast("cpu", bfloat16)`` —
what Lightning's ``precision: bf16-mixed`` (trainer.py:450) does to losses.py:195.  Run in the build
container only:

    python tests/golden/make_golden_autocast_d384.py

Stores the inputs, every loss and dL/dquery under autocast, AND the same in fp32 (no autocast): the gap
between the two is the bf16 error of the REFERENCE ITSELF (its bmm output, its dlogits and its dq are all
rounded to bf16), the yardstick for this repository's bf16 gradient tolerance.
"""

from __future__ import annotations

import pathlib
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(pathlib.Path(__file__).parent))
from xfmr_rec import losses as ref  # noqa: E402

from make_golden import LOSS_NAMES, dense_candidates  # noqa: E402

OUT = pathlib.Path(__file__).parent


def main():
    torch.manual_seed(7)
    m, cn, d = 160, 300, 384
    sc = d ** -0.5
    q, pos, neg = torch.randn(m, d) * sc, torch.randn(m, d) * sc, torch.randn(cn, d) * sc
    neg[3] = pos[5]          # duplicate of a positive inside the pool (tie => masked)
    cand = dense_candidates(pos, neg)
    rec = {"query": q.numpy(), "pos": pos.numpy(), "neg": neg.numpy()}
    cfg = ref.LossConfig()
    for tag, autocast in (("autocast", True), ("fp32", False)):
        for loss_name, cls in zip(LOSS_NAMES, ref.LOSS_CLASSES):
            qq = q.clone().requires_grad_(True)
            with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
                loss = cls(cfg)(query_embed=qq, candidate_embed=cand)
            loss.backward()
            rec[f"{tag}/loss/{loss_name}"] = np.array(loss.item(), dtype=np.float64)
            rec[f"{tag}/dq/{loss_name}"] = qq.grad.numpy()
    np.savez_compressed(OUT / "losses_pool_autocast_bf16_d384.npz", **rec)
    # the same under a non-default margin and scale: under autocast the reference rounds target * (1 - margin)
    # and logits - that to bf16 (bf16 tensor arithmetic, losses.py:527, 541), and logits * scale (losses.py:486)
    rec2 = {"query": q.numpy(), "pos": pos.numpy(), "neg": neg.numpy(), "cfg": np.array([5.0, 0.3])}
    cfg2 = ref.LossConfig(scale=5.0, margin=0.3)
    for loss_name, cls in zip(LOSS_NAMES, ref.LOSS_CLASSES):
        qq = q.clone().requires_grad_(True)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            loss = cls(cfg2)(query_embed=qq, candidate_embed=cand)
        loss.backward()
        rec2[f"autocast/loss/{loss_name}"] = np.array(loss.item(), dtype=np.float64)
        rec2[f"autocast/dq/{loss_name}"] = qq.grad.numpy()
    np.savez_compressed(OUT / "losses_pool_autocast_bf16_d384_margin.npz", **rec2)
    for name in LOSS_NAMES:
        a, b = rec[f"autocast/dq/{name}"].astype(np.float64), rec[f"fp32/dq/{name}"].astype(np.float64)
        print(f"{name:28s} loss autocast {float(rec[f'autocast/loss/{name}']):.5f} fp32 {float(rec[f'fp32/loss/{name}']):.5f}"
              f"  reference's own bf16 gradient error: norm-wise {np.linalg.norm(a - b) / np.linalg.norm(b):.2e}, "
              f"max-abs / max {np.abs(a - b).max() / np.abs(b).max():.2e}")


if __name__ == "__main__":
    main()
