"""Small driver for ncu: full-catalog top-100 search (xr_score_topk through ExactIndex.search_batch and one
CUDA-graph replay) on one GPU, 20-200 exclusions per query.
    python profiles/run_search.py [catalog_rows] [queries] [reps]"""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch

import xfmr_rec_b200 as xr
from xfmr_rec_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
u = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
raw = torch.empty((n, 384), dtype=torch.bfloat16, device=dev)
for lo in range(0, n, 1_000_000):
    raw[lo:lo + 1_000_000] = torch.randn((min(1_000_000, n - lo), 384), generator=g, device=dev).bfloat16()
idx = xr.index.ExactIndex(xr.index.ExactIndexConfig(index_metric="cosine", dtype="bf16"), dev).set_catalog(raw)
del raw
q = torch.randn((u, 384), generator=g, device=dev)
excl = [torch.randint(0, n, (int(torch.randint(20, 201, (1,)).item()),)).tolist() for _ in range(u)]
csr = ops._csr(excl, dev)
for _ in range(reps):
    s, i = idx.search_batch(q, csr, 100, max_exclusions=200)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    s, i = idx.search_batch(q, csr, 100, max_exclusions=200)
b.record()
torch.cuda.synchronize()
print(f"N={n} U={u}: {a.elapsed_time(b) / reps:.3f} ms per batch (eager); top score {float(s[0, 0]):.4f}")
