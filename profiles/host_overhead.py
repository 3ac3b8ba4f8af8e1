"""cProfile of the host side of the train step (where the non-kernel time of a step goes)."""
import cProfile
import pathlib
import pstats
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch

import xfmr_rec_b200 as xr
from oracle import xfmr_oracle as orc

b = orc.synth_batch(27278, 128, 200, dim=384, seed=0)
dev = torch.device("cuda", 0)
emb = xr.models.ItemEmbeddings(torch.from_numpy(b["table"]), add_padding_row=False).to(dev)
idx = {k: torch.from_numpy(b[k]).to(dev) for k in ("history_item_idx", "pos_item_idx", "neg_item_idx")}
tok0 = torch.from_numpy(b["token_embeddings"]).to(dev).bfloat16()
loss_fn = xr.InfoNCELoss(xr.LossConfig())


def step():
    tok = tok0.detach().requires_grad_(True)
    out = xr.models.compute_embeds(emb, tok, idx["history_item_idx"], idx["pos_item_idx"],
                                   idx["neg_item_idx"], candidate_dtype=torch.bfloat16)
    loss = loss_fn(out["query_embed"], out["candidate_embed"])
    loss.backward()


for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host-side {1e3*(t1-t0)/200:.3f} ms/step issued, {1e3*(t2-t0)/200:.3f} ms/step incl. drain")
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
