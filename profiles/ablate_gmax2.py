"""Timing ablations of the CTA-pair retrieval scoring kernel (results are garbage by construction).
mask bits: 1 no catalog TMA, 2 no epilogue TMEM loads, 4 no gmax stores, 8 no MMAs."""
import ctypes
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch

from xfmr_rec_b200 import _native as N, ops

n, u = 10_000_000, 256
g = torch.Generator(device="cuda").manual_seed(1)
cat = torch.randn((n, 384), generator=g, device="cuda").bfloat16()
q = torch.randn((u, 384), generator=g, device="cuda").bfloat16()
lib = N.lib()
tiles = (n + 127) // 128 / 74.0
for mask in (0, 1, 2, 4, 8, 6, 9, 14, 15, 7):
    lib.xr_fused_wait_stats(mask << 8, None)
    for _ in range(2):
        ops.score_groupmax(q, cat)
    torch.cuda.synchronize()
    lib.xr_fused_profile(1)
    for _ in range(4):
        ops.score_groupmax(q, cat)
    buf = (ctypes.c_float * 512)()
    k = lib.xr_fused_profile_read(buf, 512)
    lib.xr_fused_profile(0)
    ms = sum(buf[i] for i in range(k)) / k
    print(f"mask={mask:2d}: {ms:.4f} ms  {ms * 1e6 / tiles:7.0f} ns per 128-candidate tile per CTA pair")
lib.xr_fused_wait_stats(0, None)
