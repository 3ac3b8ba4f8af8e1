"""Small driver for ncu: full-catalog top-100 search (group-max path) on one GPU.
    python profiles/run_search.py [catalog_rows] [queries]"""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch

import xfmr_rec_b200 as xr

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
u = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
idx = xr.index.ExactIndex(xr.index.ExactIndexConfig(index_metric="cosine", dtype="bf16"), dev)
idx.catalog, _ = xr.ops.normalize_rows(torch.randn((n, 384), generator=g, device=dev).bfloat16(), 1e-12,
                                       torch.bfloat16)
q = torch.randn((u, 384), generator=g, device=dev)
for _ in range(3):
    s, i = idx.search_batch(q, None, 100)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    s, i = idx.search_batch(q, None, 100)
b.record()
torch.cuda.synchronize()
print(f"N={n} U={u}: {a.elapsed_time(b) / 5:.3f} ms per batch; top score {float(s[0, 0]):.4f}")
