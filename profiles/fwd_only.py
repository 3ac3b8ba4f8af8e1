"""Experiment: fused kernel forward-only (score MMA + epilogue, no dQ MMA) vs forward+backward."""
import ctypes
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch

import xfmr_rec_b200 as xr
from xfmr_rec_b200 import _native as N, ops

m, cn = 12078, 12676
g = torch.Generator(device="cuda").manual_seed(0)
q = (torch.randn(m, 384, device="cuda", generator=g) / 19.6).bfloat16()
pos = (torch.randn(m, 384, device="cuda", generator=g) / 19.6).bfloat16()
neg = (torch.randn(cn, 384, device="cuda", generator=g) / 19.6).bfloat16()
cfg = N.XrLossConfig(1, 0, 1.0, 0.5, 1)
lib = N.lib()
for grad in (True, False):
    for _ in range(3):
        ops.fused_pool_loss(q, pos, neg, 3, cfg, want_grad=grad)
    torch.cuda.synchronize()
    lib.xr_fused_profile(1)
    for _ in range(10):
        ops.fused_pool_loss(q, pos, neg, 3, cfg, want_grad=grad)
    buf = (ctypes.c_float * 512)()
    n = lib.xr_fused_profile_read(buf, 512)
    lib.xr_fused_profile(0)
    ms = sum(buf[i] for i in range(n)) / n
    fl = (4.0 if grad else 2.0) * m * (cn + 1) * 384
    print(f"grad={grad}: {ms:.4f} ms  {fl / ms / 1e9:.0f} TFLOP/s")
