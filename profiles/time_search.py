"""Full-catalog top-100 search, filter path (xr_score_topk): per-batch time eager / CUDA-graph replay with
20-200 exclusions per query, and the per-stage breakdown (CUDA events around each C-ABI call).
    python profiles/time_search.py [catalog_rows] [queries ...]"""
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch

import xfmr_rec_b200 as xr
from xfmr_rec_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
us = [int(x) for x in sys.argv[2:]] or [1, 128, 256, 1024]
d, k = 384, 100
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
raw = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
for lo in range(0, n, 1_000_000):
    raw[lo:lo + 1_000_000] = torch.randn((min(1_000_000, n - lo), d), generator=g, device=dev).bfloat16()
idx = xr.index.ExactIndex(xr.index.ExactIndexConfig(index_metric="cosine", dtype="bf16"), dev)
idx.set_catalog(raw)
del raw
cat = idx.catalog


def timed(fn, iters):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        r = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, r


for u in us:
    q = torch.randn((u, d), generator=g, device=dev)
    excl = [torch.randint(0, n, (int(torch.randint(20, 201, (1,)).item()),)).tolist() for _ in range(u)]
    csr = ops._csr(excl, dev)
    iters = 10 if u <= 1024 else 3
    pt = {"catalog_rows": n, "queries": u, "k": k}
    for name, ex, mx in (("no_excl", None, 0), ("excl_20_200", csr, 200)):
        ms, (s, i) = timed(lambda: idx.search_batch(q, ex, k, max_exclusions=mx), iters)
        plan = idx.compile_search(u, k, max_exclusions=mx)
        msg, (ps, pi) = timed(lambda: plan(q, ex, check=False), iters)
        assert not plan.overflowed()
        assert torch.equal(pi, i) and torch.equal(ps, s)
        pt[name] = {"eager_ms": round(ms, 4), "graph_ms": round(msg, 4), "queries_per_s_graph": round(u / msg * 1e3)}
        del plan
    # stage breakdown (no exclusions)
    qn, _ = ops.normalize_rows(q, 1e-12, torch.bfloat16)
    kk = k + 28
    stride = 32
    t_g, gm = timed(lambda: ops.score_groupmax(qn, cat, stride), iters)
    t_t, th = timed(lambda: ops.kth_largest(gm, kk), iters)
    t_f, fs = timed(lambda: ops.score_filter(qn, cat, th), iters)
    t_z, _ = timed(lambda: ops.filter_finalize(fs, n, th, k, k), iters)
    t_inf, _ = timed(lambda: ops.score_filter(qn, cat, torch.full_like(th, float("inf"))), iters)
    cnt = fs.counts()
    pt["stages_ms"] = {"sample_groupmax": round(t_g, 4), "topk_threshold": round(t_t, 4),
                       "score_filter": round(t_f, 4), "finalize": round(t_z, 4),
                       "score_filter_no_survivors": round(t_inf, 4)}
    pt["survivors_per_query"] = {"mean": float(cnt.float().mean()), "max": int(cnt.max()),
                                 "spilled_to_overflow_max": int(fs.o_count.max()), "n_sub": fs.n_sub, "cap_b": fs.cap_b}
    pt["filter_TFLOPs"] = round(2.0 * u * n * d / t_f / 1e9, 1)
    pt["filter_catalog_GBs"] = round(n * d * 2.0 / t_f / 1e6, 1)
    print(json.dumps(pt), flush=True)
