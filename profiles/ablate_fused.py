"""Timing ablations of the fused kernel on the bench shape (profiling aid; ablated results are
garbage by construction).  mask bits: 1 no TMA loads, 2 no epilogue math, 4 no score MMAs,
8 no dQ MMAs."""
import ctypes
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch

from xfmr_rec_b200 import _native as N, ops

m, cn = 12078, 12676
g = torch.Generator(device="cuda").manual_seed(0)
q = (torch.randn(m, 384, device="cuda", generator=g) / 19.6).bfloat16()
pos = (torch.randn(m, 384, device="cuda", generator=g) / 19.6).bfloat16()
neg = (torch.randn(cn, 384, device="cuda", generator=g) / 19.6).bfloat16()
cfg = N.XrLossConfig(1, 0, 1.0, 0.5, 1)
lib = N.lib()
tiles = ((m + 127) // 128) * ((cn + 63) // 64) / 148.0
names = {0: "full kernel", 1: "no TMA", 2: "no epilogue math", 4: "no score MMA", 8: "no dQ MMA",
         3: "no TMA, no epi math", 5: "no TMA, no score MMA", 9: "no TMA, no dQ MMA",
         6: "no epi math, no score MMA", 16: "no epi TMEM ld/st", 18: "no epi math, no epi TMEM",
         19: "MMAs only (no TMA/epi math/TMEM)", 29: "skeleton + epi math only", 31: "barriers only", 10: "no epi math, no dQ MMA", 12: "no MMAs at all",
         13: "no TMA, no MMAs", 14: "TMA only", 7: "dQ MMA only", 11: "score MMA only"}
for grad in (True, False):
    for mask in (0, 1, 2, 16, 18, 19, 4, 8, 12, 13, 29, 31):
        if not grad and mask & 8:
            continue
        lib.xr_fused_wait_stats(mask << 8, None)
        for _ in range(2):
            ops.fused_pool_loss(q, pos, neg, 3, cfg, want_grad=grad)
        torch.cuda.synchronize()
        lib.xr_fused_profile(1)
        for _ in range(5):
            ops.fused_pool_loss(q, pos, neg, 3, cfg, want_grad=grad)
        buf = (ctypes.c_float * 512)()
        n = lib.xr_fused_profile_read(buf, 512)
        lib.xr_fused_profile(0)
        ms = sum(buf[i] for i in range(n)) / n
        print(f"grad={int(grad)} mask={mask:2d} {names[mask]:28s}: {ms:.4f} ms  "
              f"{ms * 1.965e6 / tiles:7.0f} cycles/tile @1.965GHz")
lib.xr_fused_wait_stats(0, None)
