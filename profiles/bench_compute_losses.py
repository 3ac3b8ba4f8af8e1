"""What RecommenderLightningModule.compute_losses does every training step (trainer.py:213-264):
LogitsStatistics + all seven losses, autograd edge on InfoNCE, at BASELINE configs[1]
(B=128 x L=200, in-batch shared pool, bf16).  Times evaluate_all (one tensor-core pass per logit
family + the InfoNCE forward/backward) against the seven modules + LogitsStatistics called one by
one, and the all-losses kernels alone (xr_fused_profile).
    python profiles/bench_compute_losses.py [B] [L] > profiles/compute_losses_r01.json"""
import ctypes
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch

import xfmr_rec_b200 as xr
from xfmr_rec_b200 import _native as N, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
L = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
n_items, d = 27278, 384
lens = torch.randint(1, L + 1, (B,), generator=g, device=dev)
m_a = int(lens.sum())
m = int(m_a * 0.95)
s = d ** -0.5
table = (torch.randn((n_items + 1, d), generator=g, device=dev) * s).bfloat16()
q = (torch.randn((m, d), generator=g, device=dev) * s).bfloat16()
pos = table[torch.randint(1, n_items + 1, (m,), generator=g, device=dev)]
neg = table[torch.randint(1, n_items + 1, (m_a,), generator=g, device=dev)]
cand = xr.PoolCandidates(pos, neg)
cfg = xr.LossConfig()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters


def all_in_one():
    qq = q.detach().requires_grad_(True)
    out, stats = xr.losses.evaluate_all(cfg, qq, cand)
    out["loss/InfoNCELoss"].backward()
    return out, stats


mods = [cls(cfg) for cls in xr.LOSS_CLASSES]
stat_mod = xr.LogitsStatistics(cfg)


def one_by_one():
    qq = q.detach().requires_grad_(True)
    st = stat_mod(query_embed=qq, candidate_embed=cand)
    out = {type(mm).__name__: mm(query_embed=qq, candidate_embed=cand) for mm in mods}
    out["InfoNCELoss"].backward()
    return out, st


# the same through ONE sync-free, graph-replayed sequence (PoolLossStep(monitor=True)): needs a SeqBatch
import numpy as np

from oracle import xfmr_oracle as orc

sb = orc.synth_batch(n_items, B, L, dim=d, seed=0)
emb = xr.models.ItemEmbeddings(torch.from_numpy(sb["table"]), add_padding_row=False).to(dev)
mstep = xr.PoolLossStep(emb, xr.InfoNCELoss(cfg), B, L, token_dtype=torch.bfloat16, monitor=True)
pstep = xr.PoolLossStep(emb, xr.InfoNCELoss(cfg), B, L, token_dtype=torch.bfloat16)
for st in (mstep, pstep):
    st.load(torch.from_numpy(sb["token_embeddings"]).to(dev).bfloat16(),
            *(torch.from_numpy(sb[k]).to(dev) for k in ("history_item_idx", "pos_item_idx", "neg_item_idx")))
torch.cuda.synchronize()
graph_monitor_ms = timed(lambda: mstep.run())
graph_monitor_with_dict_ms = timed(lambda: (mstep.run(), mstep.loss_dict()))
graph_plain_ms = timed(lambda: pstep.run())
m_a_g, m_g = mstep.row_counts()

lib = N.lib()
c_cfg = ops.make_cfg(cfg, logits_bf16=True)
kern = {}
for name, cos in (("all_dot", False), ("all_cos", True)):
    qq, pp, nn = q, pos, neg
    if cos:
        qq, pp, nn = (ops.normalize_rows(t, 1e-8, torch.bfloat16)[0] for t in (q, pos, neg))
    cc = ops.make_cfg(cfg, logits_bf16=not cos)
    for _ in range(3):
        ops.fused_pool_all(qq, pp, nn, cc, cos)
    lib.xr_fused_profile(1)
    for _ in range(20):
        flush.zero_()
        ops.fused_pool_all(qq, pp, nn, cc, cos)
    buf = (ctypes.c_float * 64)()
    n = lib.xr_fused_profile_read(buf, 64)
    lib.xr_fused_profile(0)
    ms = sum(buf[i] for i in range(n)) / n
    kern[name] = {"kernel_ms": ms, "TFLOP/s": 2.0 * m * m_a * d / ms / 1e9,
                  "frac_of_measured_bf16": 2.0 * m * m_a * d / ms / 1e9 / 1376.9}

res = {
    "workload": f"ML-20M-shaped, B={B} x L={L}: M={m} rows x C={m_a + 1} candidates, D=384, bf16",
    "compute_losses_evaluate_all_ms": timed(all_in_one),
    "compute_losses_modules_one_by_one_ms": timed(one_by_one),
    "graph_step_with_monitor_ms": graph_monitor_ms,
    "graph_step_with_monitor_and_host_dict_ms": graph_monitor_with_dict_ms,
    "graph_step_train_loss_only_ms": graph_plain_ms,
    "graph_step_shape": f"M={m_g} rows x C={m_a_g + 1} (the bench.py batch)",
    "kernels": kern,
    "note": "wall of the device work per call (CUDA events, L2 flushed between calls, host syncs of "
            "the reference's .item() reads included: evaluate_all has one)",
}
print(json.dumps(res, indent=1))
