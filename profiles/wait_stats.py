"""Per-barrier wait cycles of the fused kernel on the bench workload (profiling aid)."""
import ctypes
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch

import xfmr_rec_b200 as xr
from oracle import xfmr_oracle as orc

b = orc.synth_batch(27278, 128, 200, dim=384, seed=0)
dev = torch.device("cuda", 0)
emb = xr.models.ItemEmbeddings(torch.from_numpy(b["table"]), add_padding_row=False).to(dev)
idx = {k: torch.from_numpy(b[k]).to(dev) for k in ("history_item_idx", "pos_item_idx", "neg_item_idx")}
tok = torch.from_numpy(b["token_embeddings"]).to(dev).bfloat16().requires_grad_(True)
out = xr.models.compute_embeds(emb, tok, idx["history_item_idx"], idx["pos_item_idx"],
                               idx["neg_item_idx"], candidate_dtype=torch.bfloat16)
fn = xr.InfoNCELoss(xr.LossConfig())
lib = xr._native.lib()
for _ in range(3):
    fn(out["query_embed"], out["candidate_embed"])
torch.cuda.synchronize()
mask = int(sys.argv[1]) if len(sys.argv) > 1 else 0
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1   # 1 wait counters + timeline, 2 timeline only
lib.xr_fused_wait_stats(mode | (mask << 8), None)
if len(sys.argv) > 3 and sys.argv[3] == "nograd":   # forward only (the retrieval kernel's structure)
    with torch.no_grad():
        fn(out["query_embed"], out["candidate_embed"])
else:
    fn(out["query_embed"], out["candidate_embed"])
torch.cuda.synchronize()
buf = (ctypes.c_uint64 * 16)()
lib.xr_fused_wait_stats(0, buf)
names = {1: "producer: q_empty", 2: "producer: ring slot free", 3: "grad issuer: p_full (weights ready)",
         4: "score issuer: q_full", 5: "grad issuer: o_empty", 6: "score issuer: s_read (S buffer handed back)",
         7: "score issuer: ring pair loaded", 8: "epilogue warps: s_full (x16 warps)",
         9: "epilogue warps: o_full (x16 warps)", 10: "epilogue warps: w_free (x16 warps)"}
m, c = out["query_embed"].size(0), out["candidate_embed"].size(1)
print(f"M={m} C={c}; wait cycles summed over 148 CTAs (per-CTA average in parentheses)")
for t, n in names.items():
    div = 148 * (16 if t in (8, 9, 10) else 1)
    print(f"  tag {t}: {buf[t]:>14d}  ({buf[t] / div:>10.0f} cycles/CTA{'/warp' if t in (8, 9, 10) else ''})  {n}")

tl = (ctypes.c_int64 * 576)()
lib.xr_fused_timeline(tl)
base = tl[0]
print(f"ablate mask {mask}")
print("tile: score_issue_start  issued  | epi_wake  ld_done  math_done st_done p_arrive | grad_wake   (cycles rel. to tile 0 start)")
for t in range(16):
    r = [tl[t * 8 + k] - base if tl[t * 8 + k] else -1 for k in range(8)]
    print(f"{t:3d}: {r[0]:8d} {r[1]:8d} | {r[2]:8d} {r[3]:8d} {r[6]:8d} {r[7]:8d} {r[4]:8d} | {r[5]:8d}")
issued = [tl[t * 8 + 1] - base for t in range(64)]
print("score 'issued' time of tiles 0..63, deltas:", [issued[t] - issued[t - 1] for t in range(1, 64)])

k0, g0, k1, g1 = tl[512], tl[513], tl[514], tl[515]
print(f"CTA 0: score issuer entry -> exit {k1 - k0} cycles in {g1 - g0} ns = {(k1 - k0) / max(g1 - g0, 1):.3f} GHz")
print("item: q_loads_issued  first_q_full  last_tile_issued  drain_start  drain_end  grad_o_empty   (cycles rel. to entry)")
for it in range(6):
    r = [tl[520 + it * 8 + k] - k0 if tl[520 + it * 8 + k] else -1 for k in range(6)]
    print(f"{it:3d}: {r[4]:9d} {r[0]:9d} {r[1]:9d} {r[2]:9d} {r[3]:9d} {r[5]:9d}")
