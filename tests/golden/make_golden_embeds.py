"""Golden vectors for candidate construction, produced by EXECUTING the reference's own
``RecommenderModel.forward`` and ``RecommenderModel.compute_embeds`` (xfmr_rec/models.py:306-345,
366-419) — run in the build container only:

    python tests/golden/make_golden_embeds.py

``xfmr_rec/models.py`` cannot be imported here (sentence_transformers is absent), so the two method
definitions are taken from the source file where it lies (``ast``: nothing is copied into this repo)
and bound to a stand-in object that supplies exactly what they touch: ``self.embeddings`` built as
models.py:247-253 builds it (``nn.Embedding.from_pretrained(weights, freeze=True, padding_idx=0)`` with
the zero padding row prepended), ``self.model`` = a stub encoder that returns fixed token embeddings
in the SentenceTransformer feature-dict form, ``self.max_seq_length``, ``self.device``,
``self.config.is_normalized``.  Inputs + the reference's outputs are stored as small ``.npz`` files.
"""

from __future__ import annotations

import ast
import pathlib
import types

import numpy as np
import torch

SRC = pathlib.Path("/root/reference/xfmr_rec/models.py")
OUT = pathlib.Path(__file__).parent


def reference_methods():
    tree = ast.parse(SRC.read_text())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "RecommenderModel")
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ("forward", "compute_embeds")]
    mod = ast.Module(body=fns, type_ignores=[])
    ns = {"torch": torch, "torch_fn": torch.nn.functional}
    exec(compile(mod, str(SRC), "exec"), ns)   # the reference's code, executed from its own file
    return ns["forward"], ns["compute_embeds"]


class StubEncoder(torch.nn.Module):
    """Stands in for the SentenceTransformer: passes the features through and adds fixed
    ``token_embeddings`` (what the BERT decoder would output), models.py:345."""

    def __init__(self, tokens):
        super().__init__()
        self.tokens = tokens

    def forward(self, features):
        l = features["inputs_embeds"].size(1)
        return {**features, "token_embeddings": self.tokens[:, -l:, :]}


def make_case(name, n_items, batch, seq_len, dim, max_seq_length, is_normalized, seed):
    g = torch.Generator().manual_seed(seed)
    weights = torch.randn((n_items, dim), generator=g) / dim ** 0.5
    weights = torch.cat([torch.zeros(1, dim), weights])                         # models.py:250
    lens = torch.randint(1, seq_len + 1, (batch,), generator=g)
    valid = torch.arange(seq_len)[None, :] < lens[:, None]
    hist = torch.randint(1, n_items + 1, (batch, seq_len), generator=g) * valid
    pos = torch.randint(1, n_items + 1, (batch, seq_len), generator=g) * valid
    pos = pos * (torch.rand((batch, seq_len), generator=g) > 0.2)               # data.py:710-721
    neg = torch.randint(1, n_items + 1, (batch, seq_len), generator=g) * valid
    tokens = torch.randn((batch, seq_len, dim), generator=g, requires_grad=True)
    fwd, compute_embeds = reference_methods()
    me = types.SimpleNamespace()
    me.embeddings = torch.nn.Embedding.from_pretrained(weights, freeze=True, padding_idx=0)  # models.py:251-253
    me.model = StubEncoder(tokens)
    me.max_seq_length = max_seq_length
    me.device = torch.device("cpu")
    me.config = types.SimpleNamespace(is_normalized=is_normalized)

    class Callable_(types.SimpleNamespace):
        def __call__(self, item_idx=None, *, item_embeds=None):
            return fwd(self, item_idx, item_embeds=item_embeds)

    me = Callable_(**me.__dict__)
    # compute_embeds indexes pos/neg with the mask of the TRUNCATED history: same length needed
    h, p, n = hist[:, -max_seq_length:], pos[:, -max_seq_length:], neg[:, -max_seq_length:]
    out = compute_embeds(me, h, p, n)
    # autograd of the query selection back to the encoder output
    w = torch.randn(out["query_embed"].shape, generator=g)
    (out["query_embed"] * w).sum().backward()
    rec = {"table": weights.numpy(), "tokens": tokens.detach().numpy()[:, -max_seq_length:],
           "history_item_idx": h.numpy(), "pos_item_idx": p.numpy(), "neg_item_idx": n.numpy(),
           "is_normalized": np.array(is_normalized), "query_embed": out["query_embed"].detach().numpy(),
           "candidate_embed": out["candidate_embed"].detach().numpy(),
           "attention_mask": out["attention_mask"].numpy(), "positive_mask": out["positive_mask"].numpy(),
           "upstream": w.numpy(), "dtokens": tokens.grad.numpy()[:, -max_seq_length:]}
    np.savez_compressed(OUT / f"embeds_{name}.npz", **rec)
    print(name, {k: v.shape for k, v in rec.items() if hasattr(v, "shape") and v.ndim})


if __name__ == "__main__":
    make_case("basic", n_items=40, batch=4, seq_len=9, dim=16, max_seq_length=32, is_normalized=False, seed=0)
    make_case("normalized", n_items=40, batch=3, seq_len=7, dim=16, max_seq_length=32, is_normalized=True, seed=1)
    make_case("truncated", n_items=60, batch=5, seq_len=12, dim=8, max_seq_length=6, is_normalized=False, seed=2)
