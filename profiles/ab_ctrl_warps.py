"""A/B: control roles (TMA producer, MMA issuers) on hardware warps 16-19 (default) vs 0-3 (round 1).
Fused train kernel at the cfg-2 shape (kernel time from the in-library CUDA events) and the retrieval
search (graph replay) at 10M x 384, U = 256.
    python profiles/ab_ctrl_warps.py [catalog_rows]"""
import ctypes
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch

import xfmr_rec_b200 as xr
from xfmr_rec_b200 import _native as N, ops

lib = N.lib()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
m, cn, d = 12078, 12677, 384
q = (torch.randn((m, d), generator=g, device=dev) / d ** 0.5).bfloat16()
pos = (torch.randn((m, d), generator=g, device=dev) / d ** 0.5).bfloat16()
neg = (torch.randn((cn, d), generator=g, device=dev) / d ** 0.5).bfloat16()
cfg = ops.make_cfg(xr.LossConfig(), logits_bf16=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000


def kernel_ms(kind, iters=20):
    for _ in range(3):
        ops.fused_pool_loss(q, pos, neg, kind, cfg)
    torch.cuda.synchronize()
    lib.xr_fused_profile(1)
    for _ in range(iters):
        flush.fill_(1)
        loss, dq, _ = ops.fused_pool_loss(q, pos, neg, kind, cfg)
    torch.cuda.synchronize()
    buf = (ctypes.c_float * 512)()
    cnt = lib.xr_fused_profile_read(buf, 512)
    lib.xr_fused_profile(0)
    ms = sorted(buf[i] for i in range(cnt))
    return ms[len(ms) // 2], float(loss.view(torch.float32)[2]), dq


out = {}
ref = None
for name, flag in (("ctrl_high", 0), ("ctrl_low", 8), ("ctrl_high_again", 0)):
    lib.xr_fused_wait_stats(flag, None)
    row = {}
    for kind_name in ("InfoNCELoss", "PairwiseLogisticLoss"):
        ms, loss, dq = kernel_ms(N.LOSS_KIND[kind_name])
        row[kind_name] = {"kernel_ms": round(ms, 4), "TFLOPs": round(4.0 * m * (cn + 1) * d / ms / 1e9, 1), "loss": loss}
        if kind_name == "InfoNCELoss":
            if ref is None:
                ref = dq.clone()
            else:
                assert torch.equal(ref, dq), "warp placement changed the result"
    out[name] = row
    print(json.dumps({name: row}), flush=True)
lib.xr_fused_wait_stats(0, None)

raw = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
for lo in range(0, n, 1_000_000):
    raw[lo:lo + 1_000_000] = torch.randn((min(1_000_000, n - lo), d), generator=g, device=dev).bfloat16()
idx = xr.index.ExactIndex(xr.index.ExactIndexConfig(index_metric="cosine", dtype="bf16"), dev).set_catalog(raw)
del raw
for u in (128, 256):
    qq = torch.randn((u, d), generator=g, device=dev)
    for name, flag in (("ctrl_high", 0), ("ctrl_low", 8)):
        lib.xr_fused_wait_stats(flag, None)
        plan = idx.compile_search(u, 100)
        for _ in range(3):
            plan(qq, check=False)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            plan(qq, check=False)
        b.record()
        torch.cuda.synchronize()
        print(json.dumps({"search": name, "queries": u, "catalog_rows": n, "graph_ms": round(a.elapsed_time(b) / 10, 4)}), flush=True)
        del plan
lib.xr_fused_wait_stats(0, None)
