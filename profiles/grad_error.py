"""bf16 gradient error, measured honestly.  Three yardsticks, all against the float64
oracle evaluated on bf16-rounded inputs AND bf16-rounded logits (what Lightning's bf16-mixed autocast does
to losses.py:195):
  (a) the REFERENCE's own autocast gradient (tests/golden/losses_pool_autocast_bf16_d384.npz),
  (b) this repository's fused tcgen05 kernel on the same inputs,
  (c) the fused kernel at the full configs[1] size (M ~ 12k x C ~ 12.7k) and one configs[4] point.
    python profiles/grad_error.py > profiles/grad_error_r02.json"""
import json
import pathlib
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import numpy as np
import torch

import xfmr_rec_b200 as xr
from oracle import xfmr_oracle as orc


def errs(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return {"normwise": float(np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30)),
            "maxabs_over_max": float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))}


def oracle_big(name, q, p, n, cfg):
    """lean_loss with BLAS logits (exact_ties=False): float64 on the bf16-rounded operands."""
    q, p, n = (orc.round_bf16(x.astype(np.float32)).astype(np.float64) for x in (q, p, n))
    logits = orc.lean_logits(q, p, n, exact_ties=False)
    logits = orc.round_bf16(logits.astype(np.float32)).astype(np.float64)
    tgt = np.zeros(q.shape[0], np.int64)
    mask = orc.mask_false_negatives(logits, tgt, cfg)
    loss, g = orc.loss_from_logits(name, logits, tgt, mask, cfg, with_grad=True)
    return loss, g[:, :1] * p + g[:, 1:] @ n


def ours(name, q, p, n, **kw):
    qt = torch.from_numpy(q).cuda().bfloat16().requires_grad_(True)
    cand = xr.PoolCandidates(torch.from_numpy(p).cuda().bfloat16(), torch.from_numpy(n).cuda().bfloat16())
    loss = getattr(xr, name)(xr.LossConfig(**kw))(qt, cand)
    ours_dq = torch.autograd.grad(loss, qt)[0]        # bf16 (the query's dtype) ...
    # ... the kernel's own fp32 gradient, before the cast to the query dtype:
    _, dq32, _ = xr.ops.fused_pool_loss(qt.detach(), cand.pos, cand.neg, xr._native.LOSS_KIND[name],
                                        xr.ops.make_cfg(xr.LossConfig(**kw), logits_bf16=True))
    return float(loss), dq32.cpu().numpy(), ours_dq.float().cpu().numpy()


out = {"golden_d384": {}, "full_size": {}}
z = np.load(ROOT / "tests" / "golden" / "losses_pool_autocast_bf16_d384.npz")
q, p, n = z["query"], z["pos"], z["neg"]
for name in ("InfoNCELoss", "NCELoss", "PairwiseHingeLoss", "PairwiseLogisticLoss"):
    want_loss, want_dq, _, _ = orc.lean_loss(name, q, p, n, orc.Config(), with_grad=True, logits_dtype="bf16")
    loss, dq32, dq16 = ours(name, q, p, n)
    out["golden_d384"][name] = {
        "loss_rel_vs_oracle": abs(loss - want_loss) / abs(want_loss),
        "loss_rel_vs_reference_autocast": abs(loss - float(z[f"autocast/loss/{name}"])) / abs(want_loss),
        "reference_autocast_vs_oracle": errs(z[f"autocast/dq/{name}"], want_dq),
        "fused_fp32_grad_vs_oracle": errs(dq32, want_dq),
        "fused_bf16_cast_grad_vs_oracle": errs(dq16, want_dq),
        "fused_vs_reference_autocast": errs(dq32, z[f"autocast/dq/{name}"])}

for tag, (m, cn) in {"configs[1] 12078 x 12677": (12078, 12677), "configs[4] 2048 x 100000": (2048, 100_000)}.items():
    rng = np.random.default_rng(m)
    q = (rng.standard_normal((m, 384)) / 384 ** 0.5).astype(np.float32)
    p = (rng.standard_normal((m, 384)) / 384 ** 0.5).astype(np.float32)
    n = (rng.standard_normal((cn, 384)) / 384 ** 0.5).astype(np.float32)
    n[: min(m, cn) // 2] = p[: min(m, cn) // 2]          # in-batch: half of the positives sit in the pool
    for name, kw in (("InfoNCELoss", {}), ("PairwiseLogisticLoss", {"margin": 0.0})):
        t0 = time.time()
        want_loss, want_dq = oracle_big(name, q, p, n, orc.Config(**kw))
        t_or = time.time() - t0
        loss, dq32, dq16 = ours(name, q, p, n, **kw)
        out["full_size"][f"{tag} {name}"] = {
            "loss_rel_vs_oracle": abs(loss - want_loss) / abs(want_loss),
            "fused_fp32_grad_vs_oracle": errs(dq32, want_dq), "oracle_seconds": round(t_or, 1)}
print(json.dumps(out, indent=1))
