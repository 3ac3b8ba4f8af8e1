"""Small driver for ncu: a few sync-free train steps (xr_pool_step, issued without graph capture so
every kernel shows up as its own launch) of the bench workload.
    python profiles/run_pool_step.py [steps] [batch] [--monitor]"""
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
step = xr.PoolLossStep(emb, xr.InfoNCELoss(xr.LossConfig()), batch_size, 200, use_graph=False,
                       monitor="--monitor" in sys.argv)
tok = torch.from_numpy(b["token_embeddings"]).to(dev).bfloat16()
idx = [torch.from_numpy(b[k]).to(dev) for k in ("history_item_idx", "pos_item_idx", "neg_item_idx")]
for _ in range(steps):
    loss, dtok = step(tok, *idx)
torch.cuda.synchronize()
print("loss", float(loss), "counts", step.row_counts())
