"""SeqBatch construction throughput: xr_seq_sample_batch on the device vs the
reference-style per-example Python (oracle.seq_example_reference_style, data.py:669-785) on one
host core (the reference runs it in DataLoader workers, data.py:905-927, num_workers=1 default).
ML-20M-shaped synthetic histories (27,278 items; history lengths ~ lognormal, mean ~145, max 2,000),
max_seq_length 200.
    python profiles/bench_seqbatch.py > profiles/seqbatch_r01.json"""
import json
import pathlib
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import numpy as np
import torch

import xfmr_rec_b200 as xr
from oracle import xfmr_oracle as orc

n_items, n_users, L = 27278, 20000, 200
rng = np.random.default_rng(0)
lens = np.clip(rng.lognormal(4.4, 1.0, n_users).astype(np.int64), 20, 2000)
hs = [rng.integers(1, n_items + 1, size=int(n)).astype(np.int64) for n in lens]
ls = [np.concatenate([rng.random(int(n) - 1) < 0.6, [True]]) for n in lens]
s = xr.data.SeqBatchSampler(xr.data.SeqDataConfig(L, 0), hs, ls, n_items, device="cuda", seed=1)
res = {"workload": f"{n_users} users, {int(lens.sum())} events (mean {lens.mean():.0f}, max {lens.max()}), "
                   f"{n_items} items, max_seq_length {L}, rows {len(s)}", "device": []}
g = torch.Generator(device="cuda").manual_seed(0)
for B in (128, 1024, 8192):
    rows = torch.randint(0, len(s), (B,), generator=g, device="cuda")
    for i in range(3):
        out = s.sample(rows, step=i, return_lengths=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 50
    a.record()
    for i in range(iters):
        out = s.sample(rows, step=10 + i, return_lengths=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    pos_filled = int((out["history_item_idx"] != 0).sum())
    res["device"].append({"batch": B, "ms_per_batch": ms, "seq_per_s": B / ms * 1e3,
                          "positions_per_s": pos_filled / ms * 1e3,
                          "output_GB/s": 3 * B * L * 8 / ms / 1e6})
# CPU reference-style, one core
all_idx = set(range(1, n_items + 1))
r = np.random.default_rng(0)
rh = s.row_hist.cpu().numpy()
pick = rh[np.random.default_rng(1).integers(0, len(s), 200)]
t0 = time.perf_counter()
for h in pick:
    orc.seq_example_reference_style(r, hs[h], ls[h], all_idx, L, 0)
dt = time.perf_counter() - t0
res["cpu_reference_style"] = {"examples": len(pick), "seq_per_s": len(pick) / dt, "cores": 1,
                              "ms_per_example": dt / len(pick) * 1e3}
res["speedup_b1024_vs_one_core"] = res["device"][1]["seq_per_s"] / res["cpu_reference_style"]["seq_per_s"]
print(json.dumps(res, indent=1))
