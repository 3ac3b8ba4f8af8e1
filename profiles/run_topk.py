"""ncu driver: top-k over a materialised (64, 10M) fp32 score matrix."""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch

import xfmr_rec_b200 as xr

u, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 10_000_000)
s = torch.randn(u, n, device="cuda")
for _ in range(3):
    a, b = xr.ops.topk(s, 100)
torch.cuda.synchronize()
print(a[0, :3].tolist(), b[0, :3].tolist())
