"""The dense (M, C, D) candidate tensor of the reference's own API (models.py:408-416 / losses.py:195): a
batched GEMV at 0.5-1 FLOP/B, graded on HBM.  Logits pass, dL/dq pass and the module step
(InfoNCE forward + backward = two passes over the M*C*D bytes).
    python profiles/bench_dense.py > profiles/dense_r01.jsonl"""
import sys, json
import pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch
import xfmr_rec_b200 as xr
from xfmr_rec_b200 import ops, _native as N
dev = torch.device("cuda")
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for dt, bs in ((torch.float32, 4), (torch.bfloat16, 2)):
    for m, c in ((1024, 1025), (2048, 513)):
        d = 384
        q = torch.randn(m, d, device=dev).to(dt)
        cand = torch.randn(m, c, d, device=dev).to(dt)
        byts = m * c * d * bs
        ms_l = timeit(lambda: ops.logits_dense(q, cand, None, cosine=False))
        logits, _ = ops.logits_dense(q, cand, None, cosine=False)
        cfg = ops.make_cfg(xr.LossConfig())
        _, _, dl = ops.rowloss(logits, c, cfg, N.TARGET_FIRST, None, N.LOSS_KIND["InfoNCELoss"])
        ms_g = timeit(lambda: ops.dq_dense(dl, q, cand, cosine=False))
        fn = xr.InfoNCELoss(xr.LossConfig())
        def step():
            qq = q.detach().requires_grad_(True)
            fn(qq, cand).backward()
        ms_s = timeit(step)
        print(json.dumps({"dtype": str(dt), "M": m, "C": c, "GB": byts / 1e9, "logits_ms": ms_l, "logits_GB/s": byts / ms_l / 1e6,
                          "dq_ms": ms_g, "dq_GB/s": byts / ms_g / 1e6, "step_ms": ms_s, "step_GB/s(2 passes)": 2 * byts / ms_s / 1e6}))
