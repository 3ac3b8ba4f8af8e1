"""BASELINE config 3 on one B200: AlignmentContrastiveLoss (CCL, cosine) with 512 sampled negatives
per positive (C = 513), ML-32M-shaped table (87,585 items x 384), 6,400 rows per GPU
(= global batch 1,024 sequences x L=50 over 8 GPUs) and the 1,024-row reading.
    python profiles/bench_cfg3.py [--json out.json]
Per kernel: CUDA events on the launching stream, algorithmic bytes M*C*D*b per pass against the
measured HBM copy bandwidth (the 67 MB bf16 / 134 MB fp32 table is largely L2-resident, so the
fraction can exceed 1: it is reported against HBM because that is the roofline SURVEY 8d names)."""
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch

import xfmr_rec_b200 as xr
from xfmr_rec_b200 import _native as N, ops

peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
HBM = peaks.get("hbm_gbs", 6650.0)
dev = torch.device("cuda", 0)
N_ITEMS, D, K = 87585, 384, 512
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps


g = torch.Generator(device=dev).manual_seed(0)
table32 = torch.randn(N_ITEMS + 1, D, device=dev, generator=g) / D ** 0.5
table32[0] = 0
out = []
for dt, bsz in ((torch.bfloat16, 2), (torch.float32, 4)):
    table = table32.to(dt)
    _, table_inv = ops.normalize_rows(table, 1e-8, want_y=False)
    for m in (6400, 1024):
        q = (torch.randn(m, D, device=dev, generator=g) / D ** 0.5).to(dt)
        idx = torch.randint(1, N_ITEMS + 1, (m, K + 1), device=dev, generator=g)
        cand = xr.SampledCandidates(table, idx, table_inv)
        for name in ("AlignmentContrastiveLoss", "InfoNCELoss"):
            loss_fn = getattr(xr, name)(xr.LossConfig())

            def step():
                qq = q.detach().requires_grad_(True)
                loss = loss_fn(qq, cand)
                loss.backward()
                return loss

            ms = timeit(step)
            byts = m * (K + 1) * D * bsz
            # the pieces
            cos = name == "AlignmentContrastiveLoss"
            q_inv = ops.normalize_rows(q, 1e-8, want_y=False)[1] if cos else None
            t_inv = table_inv if cos else None
            ms_logits = timeit(lambda: ops.logits_sampled(q, table, idx, t_inv, q_inv))
            logits = ops.logits_sampled(q, table, idx, t_inv, q_inv)
            cfg = ops.make_cfg(xr.LossConfig())
            kind = N.LOSS_KIND[name]
            ms_rowloss = timeit(lambda: ops.rowloss(logits, K + 1, cfg, N.TARGET_FIRST, None, kind))
            _, _, dl = ops.rowloss(logits, K + 1, cfg, N.TARGET_FIRST, None, kind)
            ms_dq = timeit(lambda: ops.dq_sampled(dl, q, table, idx, t_inv, q_inv))
            nz = float((dl != 0).float().mean())
            ms_one = timeit(lambda: ops.sampled_step(q, table, idx, cfg, t_inv, q_inv, kind))
            ms_one_fwd = timeit(lambda: ops.sampled_step(q, table, idx, cfg, t_inv, q_inv, -1))
            xr.losses._SAMPLED_ONE_PASS = False
            ms_three = timeit(step)
            xr.losses._SAMPLED_ONE_PASS = True
            rec = {"config": "cfg3", "loss": name, "dtype": str(dt).replace("torch.", ""), "rows_M": m, "C": K + 1,
                   "step_ms": ms, "rows_per_s": m / ms * 1e3,
                   "logits_ms": ms_logits, "logits_GB/s": byts / ms_logits / 1e6,
                   "logits_frac_of_measured_hbm": byts / ms_logits / 1e6 / HBM,
                   "one_pass_kernel_ms": ms_one, "one_pass_forward_only_ms": ms_one_fwd,
                   "one_pass_GB/s": byts * (1 + nz) / ms_one / 1e6,
                   "one_pass_frac_of_measured_hbm": byts * (1 + nz) / ms_one / 1e6 / HBM,
                   "step_ms_three_launches": ms_three,
                   "rowloss_ms": ms_rowloss, "dq_ms": ms_dq, "dq_nonzero_weight_frac": nz,
                   "dq_GB/s_touched": byts * nz / ms_dq / 1e6}
            out.append(rec)
            print(json.dumps(rec))
if "--json" in sys.argv:
    pathlib.Path(sys.argv[sys.argv.index("--json") + 1]).write_text(json.dumps(out, indent=1))
