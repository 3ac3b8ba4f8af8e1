"""BASELINE config 5: BPR (PairwiseLogisticLoss, margin 0) and SSM (InfoNCELoss) sweep,
batch 256-8192 x 1k-1M shared candidates, fused tcgen05 epilogue vs the reference's logits
materialisation (Q @ C^T -> (B, N) tensor -> the reference's loss arithmetic, forward + backward
w.r.t. Q) written in plain torch ops — both on the same B200, bf16 operands.
    python profiles/sweep_cfg5.py [--json out.json] [--quick]
TFLOP/s = 4*B*N*D / time (scores + dQ)."""
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch
import torch.nn.functional as F

import xfmr_rec_b200 as xr

peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
PEAK = peaks.get("bf16_tflops_sustained", 1400.0)
dev = torch.device("cuda", 0)
D = 384
quick = "--quick" in sys.argv


def timeit(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def materialised(name, q, pos, neg):
    """losses.py:195 (bf16 bmm under autocast) + :283-292 mask + :479-488 / :536-543, lean form:
    the (B, 1+N) logits tensor exists in HBM, the (B, 1+N, D) candidate tensor does not."""
    q = q.detach().requires_grad_(True)
    pl = (q * pos).sum(-1, keepdim=True)                       # bf16, (B, 1)
    nl = q @ neg.T                                              # bf16, (B, N)
    mask = nl < pl                                              # losses.py:292
    if name == "InfoNCELoss":
        logits = torch.cat([pl, nl.masked_fill(~mask, float("-inf"))], dim=1).float()
        loss = F.cross_entropy(logits, torch.zeros(q.size(0), dtype=torch.long, device=q.device),
                               reduction="sum")
    else:                                                       # PairwiseLogisticLoss, margin 0
        x = F.softplus((nl - pl).float())
        w = mask.float()
        loss = ((x * w).sum(-1) / (w.sum(-1) + 1e-9)).sum()
    loss.backward()
    return loss


out = []
g = torch.Generator(device=dev).manual_seed(0)
batches = (256, 1024, 8192) if quick else (256, 512, 1024, 2048, 4096, 8192)
cands = (1000, 100_000) if quick else (1000, 10_000, 100_000, 1_000_000)
for name, margin in (("InfoNCELoss", 0.5), ("PairwiseLogisticLoss", 0.0)):
    loss_fn = getattr(xr, name)(xr.LossConfig(margin=margin))
    for n in cands:
        neg = (torch.randn(n, D, device=dev, generator=g) / D ** 0.5).bfloat16()
        for b in batches:
            q = (torch.randn(b, D, device=dev, generator=g) / D ** 0.5).bfloat16()
            pos = (torch.randn(b, D, device=dev, generator=g) / D ** 0.5).bfloat16()
            cand = xr.PoolCandidates(pos, neg)

            def fused():
                qq = q.detach().requires_grad_(True)
                loss = loss_fn(qq, cand)
                loss.backward()
                return loss

            flops = 4.0 * b * (n + 1) * D
            reps = 3 if flops > 2e12 else 10
            ms_f = timeit(fused, reps)
            rec = {"config": "cfg5", "loss": name, "B": b, "N": n, "fused_ms": ms_f,
                   "fused_TFLOP/s": flops / ms_f / 1e9, "fused_frac_of_measured_bf16": flops / ms_f / 1e9 / PEAK}
            if b * n * 4 <= (8 << 30):   # the materialised path needs ~4 x B x N x 4 bytes
                try:
                    ms_m = timeit(lambda: materialised(name, q, pos, neg), reps)
                    lf, lm = float(fused()), float(materialised(name, q, pos, neg))
                    rec.update({"materialised_ms": ms_m, "speedup": ms_m / ms_f,
                                "logits_bytes_avoided": b * (n + 1) * 4,
                                "loss_rel_diff": abs(lf - lm) / max(abs(lm), 1e-30)})
                except torch.OutOfMemoryError:
                    rec["materialised_ms"] = None
                torch.cuda.empty_cache()
            out.append(rec)
            print(json.dumps(rec), flush=True)
        del neg
if "--json" in sys.argv:
    pathlib.Path(sys.argv[sys.argv.index("--json") + 1]).write_text(json.dumps(out, indent=1))
