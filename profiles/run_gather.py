"""ncu driver: gather of 1M rows from a 3 GB bf16 item table (HBM-resident, larger than L2)."""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch

import xfmr_rec_b200 as xr

n_rows, rows = 4_000_000, 1_000_000
table = torch.randn((n_rows, 384), device="cuda").bfloat16()
idx = torch.randint(0, n_rows, (rows,), device="cuda")
for _ in range(3):
    out = xr.ops.gather_rows(table, idx)
torch.cuda.synchronize()
print(out.shape, float(out[0, 0]))
