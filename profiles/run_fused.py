"""Small driver for ncu: the fused train kernel (InfoNCE, bf16-rounded logits) at the configs[1] shape.
    python profiles/run_fused.py [iters]"""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch

import xfmr_rec_b200 as xr
from xfmr_rec_b200 import _native as N, ops

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
m, cn, d = 12078, 12677, 384
q = (torch.randn((m, d), generator=g, device=dev) / d ** 0.5).bfloat16()
pos = (torch.randn((m, d), generator=g, device=dev) / d ** 0.5).bfloat16()
neg = (torch.randn((cn, d), generator=g, device=dev) / d ** 0.5).bfloat16()
cfg = ops.make_cfg(xr.LossConfig(), logits_bf16=True)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    loss, dq, _ = ops.fused_pool_loss(q, pos, neg, N.LOSS_KIND["InfoNCELoss"], cfg)
torch.cuda.synchronize()
print("loss", float(loss.view(torch.float32)[2]))
