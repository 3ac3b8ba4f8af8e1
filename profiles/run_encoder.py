"""Small driver for ncu: the train step with the sequence encoder (B = 128 x L = 200, 2 layers, bf16-mixed).
    python profiles/run_encoder.py [reps]"""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch

import xfmr_rec_b200 as xr
from xfmr_rec_b200.data import synthetic_batch
from xfmr_rec_b200.encoder import EncoderConfig, SeqEncoder, encoder_train_step

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
B, L = 128, 200
b = synthetic_batch(27278, B, L, dim=384, seed=0)
table = torch.from_numpy(b["table"]).to(dev)
hist, pos, neg = (torch.from_numpy(b[k]).to(dev) for k in ("history_item_idx", "pos_item_idx", "neg_item_idx"))
enc = SeqEncoder(EncoderConfig(num_hidden_layers=2, intermediate_size=1536, max_seq_length=L),
                 compute_dtype=torch.bfloat16).to(dev)
emb = xr.models.ItemEmbeddings(table, add_padding_row=False).to(dev)
step = xr.PoolLossStep(emb, xr.InfoNCELoss(xr.LossConfig()), B, L, token_dtype=torch.float32, logits_bf16=True)
for _ in range(reps):
    enc.zero_grad(set_to_none=True)
    out = encoder_train_step(enc, step, table, hist, pos, neg)
torch.cuda.synchronize()
print("ok", float(out["loss"]) if isinstance(out, dict) else out)
