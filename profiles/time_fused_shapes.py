"""Fused InfoNCE forward+backward kernel alone at several (rows, pool) shapes: CUDA events inside the library
around each launch (xr_fused_profile), L2 flushed between launches.  Shapes: configs[1] at B = 128 and at
B = 64 (the config does not fix B), and a few stream-K corner shapes.
    python profiles/time_fused_shapes.py > profiles/fused_shapes_r02.json"""
import ctypes as C
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch

import xfmr_rec_b200 as xr
from xfmr_rec_b200 import _native as N, ops

peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
cfg = ops.make_cfg(xr.LossConfig(), logits_bf16=True)
lib = N.lib()
out = []
for (m, cn, tag) in [(12078, 12677, "configs[1] B=128"), (5968, 6316, "configs[1] B=64"), (3000, 3100, "B=32"),
                     (24000, 25000, "B=256"), (128, 100000, "one row block x 100k"), (8192, 100000, "cfg5 8192 x 100k")]:
    q = (torch.randn((m, 384), generator=g, device=dev) / 384 ** 0.5).bfloat16()
    pos = (torch.randn((m, 384), generator=g, device=dev) / 384 ** 0.5).bfloat16()
    neg = (torch.randn((cn, 384), generator=g, device=dev) / 384 ** 0.5).bfloat16()
    for _ in range(3):
        ops.fused_pool_loss(q, pos, neg, N.LOSS_KIND["InfoNCELoss"], cfg)
    torch.cuda.synchronize()
    lib.xr_fused_profile(1)
    for _ in range(10):
        flush.fill_(1)
        ops.fused_pool_loss(q, pos, neg, N.LOSS_KIND["InfoNCELoss"], cfg)
    torch.cuda.synchronize()
    buf = (C.c_float * 64)()
    n = lib.xr_fused_profile_read(buf, 64)
    lib.xr_fused_profile(0)
    ms = sorted(buf[i] for i in range(n))
    med = ms[len(ms) // 2]
    tf = 4.0 * m * cn * 384 / med / 1e9
    out.append({"shape": tag, "rows_M": m, "candidates_C": cn, "kernel_ms": med, "kernel_ms_min": ms[0], "TFLOP/s": tf,
                "frac_of_burst_peak": tf / peaks["bf16_tflops"], "frac_of_sustained_peak": tf / peaks["bf16_tflops_sustained"]})
print(json.dumps(out, indent=1))
