"""Golden vectors for the sequence encoder by EXECUTING the reference's encoder class:
``transformers.models.bert.BertModel`` built exactly as ``xfmr_rec/models.py:init_bert`` builds it
(``BertConfig(vocab_size, hidden_size, num_hidden_layers, num_attention_heads, intermediate_size,
max_position_embeddings, is_decoder=True)``, models.py:92-101) and called as ``RecommenderModel.forward`` calls it
(``inputs_embeds = table[item_idx]``, ``attention_mask = (inputs_embeds != 0).any(-1)``, models.py:336-345),
in eval mode (no dropout), fp32, on the CPU.  Run in the build container only:

    python tests/golden/make_golden_encoder.py

To keep the fixture small the weights are NOT stored: both this script and the tests fill every parameter
from ``torch.Generator().manual_seed(seed)`` in ``sorted(state_dict)`` order (tests/test_encoder.py:
``seeded_state_dict``).  Stored: inputs, token embeddings, mean-pooled sentence embeddings, and the
gradients of every 1-D parameter and of the position embeddings in full, of every weight matrix as a
strided sample (rows ::8, columns ::8) plus its Frobenius norm.
"""

from __future__ import annotations

import pathlib

import numpy as np
import torch
from transformers.models.bert import BertConfig, BertModel

OUT = pathlib.Path(__file__).parent


def seeded_state_dict(ref_sd: dict, seed: int) -> dict:
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(ref_sd):
        shape = tuple(ref_sd[k].shape)
        if k.endswith("LayerNorm.weight"):
            out[k] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            out[k] = 0.05 * torch.randn(shape, generator=g)
    return out


def run(tag, *, layers, inter, max_pos, batch, seq_len, n_items, seed):
    cfg = BertConfig(vocab_size=1, hidden_size=384, num_hidden_layers=layers, num_attention_heads=12,
                     intermediate_size=inter, max_position_embeddings=max_pos, is_decoder=True)
    model = BertModel(cfg).eval()
    sd = seeded_state_dict({k: v for k, v in model.state_dict().items() if v.dtype.is_floating_point}, seed)
    model.load_state_dict(sd, strict=False)
    g = torch.Generator().manual_seed(seed + 1)
    table = torch.randn((n_items + 1, 384), generator=g) / 384 ** 0.5
    table[0] = 0.0
    lens = torch.randint(1, seq_len + 1, (batch,), generator=g)
    lens[0] = seq_len
    idx = torch.randint(1, n_items + 1, (batch, seq_len), generator=g)
    idx = idx * (torch.arange(seq_len)[None, :] < lens[:, None])          # right padding with 0 (data.py:801)
    inputs_embeds = table[idx]                                            # models.py:336-338
    attention_mask = (inputs_embeds != 0).any(-1).long()                  # models.py:343
    out = model(inputs_embeds=inputs_embeds, attention_mask=attention_mask).last_hidden_state
    m = attention_mask[..., None].float()
    sent = (out * m).sum(1) / m.sum(1).clamp(min=1e-9)                    # Pooling(mean)
    upstream = torch.randn(out.shape, generator=g) * attention_mask[..., None]   # padded rows get no gradient
    model.zero_grad()
    (out * upstream).sum().backward()
    rec = {"idx": idx.numpy(), "table": table.numpy(), "token_embeddings": out.detach().numpy(),
           "sentence_embedding": sent.detach().numpy(), "upstream": upstream.numpy(),
           "config": np.array([layers, inter, max_pos, seed])}
    for k, p in model.named_parameters():
        if p.grad is None:
            continue          # pooler / word embeddings: not on the path
        gr = p.grad.detach()
        if gr.dim() == 1 or "position_embeddings" in k or "token_type" in k:
            rec["grad/" + k] = gr.numpy()
        else:
            rec["gradsub/" + k] = gr[::8, ::8].contiguous().numpy()
            rec["gradnorm/" + k] = np.array(float(gr.norm()))
    np.savez_compressed(OUT / f"encoder_{tag}.npz", **rec)
    print(tag, "tokens", tuple(out.shape), "grad keys", sum(k.startswith("grad") for k in rec))


if __name__ == "__main__":
    torch.manual_seed(0)
    # the reference's default topology (models.py:39-45): 1 layer, 12 heads, intermediate 48, max length 32
    run("default_1layer", layers=1, inter=48, max_pos=32, batch=4, seq_len=20, n_items=60, seed=11)
    # BASELINE configs[0]: "2-layer d=384 encoder"
    run("2layer_i128", layers=2, inter=128, max_pos=24, batch=3, seq_len=24, n_items=80, seed=23)
