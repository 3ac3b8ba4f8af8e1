"""Small driver for ncu: a few train steps of the bench workload (no CPU baseline, no retrieval).
    python profiles/run_step.py [steps] [batch]"""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch

import xfmr_rec_b200 as xr
from oracle import xfmr_oracle as orc

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
batch_size = int(sys.argv[2]) if len(sys.argv) > 2 else 128
b = orc.synth_batch(27278, batch_size, 200, dim=384, seed=0)
dev = torch.device("cuda", 0)
emb = xr.models.ItemEmbeddings(torch.from_numpy(b["table"]), add_padding_row=False).to(dev)
idx = {k: torch.from_numpy(b[k]).to(dev) for k in ("history_item_idx", "pos_item_idx", "neg_item_idx")}
tok0 = torch.from_numpy(b["token_embeddings"]).to(dev).bfloat16()
loss_fn = xr.InfoNCELoss(xr.LossConfig())
for _ in range(steps):
    tok = tok0.detach().requires_grad_(True)
    out = xr.models.compute_embeds(emb, tok, idx["history_item_idx"], idx["pos_item_idx"],
                                   idx["neg_item_idx"], candidate_dtype=torch.bfloat16)
    loss = loss_fn(out["query_embed"], out["candidate_embed"])
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss.detach()), "M", out["query_embed"].size(0), "C", out["candidate_embed"].size(1))
