"""Sharded retrieval exchange: NVLink peer-memory merge (PeerExchange / xr_topk_merge_peers) against
the NCCL all-gather path and against the unsharded search.  Launch under torchrun on >= 2 GPUs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        profiles/run_peer_exchange.py [catalog_rows] [queries]
Exit code 0 and a JSON line from rank 0 when both exchanges give the unsharded result bit for bit."""
import json
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch
import torch.distributed as dist

import xfmr_rec_b200 as xr
from xfmr_rec_b200.dist import ShardedIndex, shard_range

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
u = int(sys.argv[2]) if len(sys.argv) > 2 else 256
k = 100
rank, world, local = (int(os.environ[x]) for x in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator(device=dev).manual_seed(7)            # the SAME catalog on every rank
cat = torch.randn((n, 384), generator=g, device=dev).bfloat16()
cat[123] = cat[77]                                        # a tie across / inside shards
cat[n - 5] = cat[77]
q = torch.randn((u, 384), generator=g, device=dev)
excl = [[77] if r % 3 == 0 else [] for r in range(u)]
cfg = xr.index.ExactIndexConfig(index_metric="cosine", dtype="bf16")
full = xr.index.ExactIndex(cfg, dev).set_catalog(cat)
want_s, want_i = full.search_batch(q, excl, k)
lo, hi = shard_range(n, rank, world)
res = {}
for mode in ("nccl", "peer"):
    idx = xr.index.ExactIndex(cfg, dev, row_offset=lo)
    idx.catalog = full.catalog[lo:hi]
    sh = ShardedIndex(idx, exchange=mode)
    for _ in range(3):
        s, i = sh.search_batch(q, excl, k)
    torch.cuda.synchronize()
    ok = bool(torch.equal(i, want_i) and torch.equal(s, want_s))
    dist.barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for a, b in evs:
        a.record()
        s, i = sh.search_batch(q, excl, k)
        b.record()
    torch.cuda.synchronize()
    ok = ok and bool(torch.equal(i, want_i) and torch.equal(s, want_s))
    in_order = [round(a.elapsed_time(b), 3) for a, b in evs]
    ms = sorted(in_order)
    t = torch.tensor([ms[len(ms) // 2], 0.0 if ok else 1.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[mode] = {"ms_per_search_median_max_over_ranks": float(t[0]), "equals_unsharded": float(t[1]) == 0.0,
                 "rank0_ms_in_order": in_order}
if rank == 0:
    print(json.dumps({"world": world, "catalog_rows": n, "queries": u, "k": k, **res}))
bad = [m for m, r in res.items() if not r["equals_unsharded"]]
dist.destroy_process_group()
sys.exit(1 if bad else 0)
