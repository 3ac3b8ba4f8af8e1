"""Sequence encoder at the configs[1] shape: B = 128 x L = 200, 2 layers, d = 384, 12 heads,
intermediate 1536 (all-MiniLM-L6-v2's), bf16-mixed.  This repository's SeqEncoder (fused kernels around cuBLAS
GEMMs) against the reference's encoder class (transformers BertModel, sdpa attention) under torch.autocast on
the same GPU; then the whole train step: encoder forward -> PoolLossStep -> encoder backward.
    python profiles/bench_encoder.py > profiles/encoder_r02.json"""
import json
import pathlib
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch

import xfmr_rec_b200 as xr
from xfmr_rec_b200.data import synthetic_batch
from xfmr_rec_b200.encoder import EncoderConfig, GraphedEncoderStep, SeqEncoder, encoder_train_step

dev = torch.device("cuda", 0)
B, L, N_ITEMS = 128, 200, 27278
b = synthetic_batch(N_ITEMS, B, L, dim=384, seed=0)
table = torch.from_numpy(b["table"]).to(dev)
hist, pos, neg = (torch.from_numpy(b[k]).to(dev) for k in ("history_item_idx", "pos_item_idx", "neg_item_idx"))
cfg = EncoderConfig(num_hidden_layers=2, intermediate_size=1536, max_seq_length=L)
out = {"shape": {"batch": B, "seq_len": L, "layers": 2, "hidden": 384, "heads": 12, "intermediate": 1536}}


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return a.elapsed_time(e) / reps


for name, cd in (("bf16", torch.bfloat16), ("fp32", torch.float32)):
    enc = SeqEncoder(cfg, compute_dtype=cd).to(dev).train()      # dropout 0.1 / 0.1 (HF defaults), as the reference trains
    up = torch.randn(B, L, 384, device=dev)

    def fwd():
        with torch.no_grad():
            return enc.encode_tokens(hist, table)[0]

    def fwd_bwd():
        enc.zero_grad(set_to_none=True)
        tok, _ = enc.encode_tokens(hist, table)
        tok.backward(up)

    out[f"ours_{name}"] = {"forward_ms": timeit(fwd), "forward_backward_ms": timeit(fwd_bwd)}
    if cd == torch.bfloat16:
        emb = xr.models.ItemEmbeddings(table, add_padding_row=False).to(dev)
        step = xr.PoolLossStep(emb, xr.InfoNCELoss(xr.LossConfig()), B, L, token_dtype=torch.float32, logits_bf16=True)

        def train():
            enc.zero_grad(set_to_none=True)
            return encoder_train_step(enc, step, table, hist, pos, neg)

        ms = timeit(train)
        out["train_step_with_encoder"] = {"ms_per_step": ms, "seq_per_s": B / ms * 1e3,
                                          "note": "encoder forward (bf16-mixed) -> sync-free scoring-and-loss step -> "
                                                  "encoder backward; parameter gradients left in .grad (no optimizer)"}
        gstep = GraphedEncoderStep(enc, xr.PoolLossStep(emb, xr.InfoNCELoss(xr.LossConfig()), B, L, token_dtype=torch.float32,
                                                        logits_bf16=True, use_graph=False), table, L)
        ms = timeit(lambda: gstep(hist, pos, neg), reps=50)
        out["train_step_with_encoder_one_graph"] = {"ms_per_step": ms, "seq_per_s": B / ms * 1e3}
        del gstep
    del enc

try:
    t0 = time.time()
    from transformers.models.bert import BertConfig, BertModel

    hf = BertModel(BertConfig(vocab_size=1, hidden_size=384, num_hidden_layers=2, num_attention_heads=12,
                              intermediate_size=1536, max_position_embeddings=L, is_decoder=True,
                              hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)).to(dev).train()
    inputs = table[hist]
    mask = (inputs != 0).any(-1).long()
    up = torch.randn(B, L, 384, device=dev)

    def hf_fwd_bwd():
        hf.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            o = hf(inputs_embeds=inputs, attention_mask=mask).last_hidden_state
        o.backward(up.to(o.dtype))

    def hf_fwd():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return hf(inputs_embeds=inputs, attention_mask=mask).last_hidden_state

    out["reference_bert_bf16_autocast_same_gpu"] = {"forward_ms": timeit(hf_fwd), "forward_backward_ms": timeit(hf_fwd_bwd),
                                                    "import_s": round(time.time() - t0, 1),
                                                    "note": "transformers BertModel (the class models.py:92-101 builds), eager, training mode "
                                                            "(dropout 0.1 / 0.1 as the reference trains), torch.autocast(bfloat16), gather done by torch"}
except Exception as e:   # transformers missing on the box
    out["reference_bert_bf16_autocast_same_gpu"] = {"unavailable": repr(e)}
print(json.dumps(out, indent=1))
