"""Golden for the whole training-step log dict, produced by EXECUTING the reference's own
``RecommenderLightningModule.compute_losses`` (xfmr_rec/trainer.py:213-264) on top of its own
``RecommenderModel.forward`` / ``compute_embeds`` (models.py:306-345, 366-419) and its own loss
classes (losses.py, importable) — build container only:

    python tests/golden/make_golden_compute_losses.py

trainer.py / models.py cannot be imported (lightning, sentence_transformers absent): the method
definitions are taken from the source files where they lie (ast; nothing is copied into this repo)
and bound to stand-in objects carrying exactly the attributes they touch.
"""

from __future__ import annotations

import ast
import json
import pathlib
import sys
import types

import numpy as np
import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent))   # make_golden_embeds lives beside this file
sys.path.insert(0, "/root/reference")
from xfmr_rec import losses as ref_losses  # noqa: E402

from make_golden_embeds import StubEncoder, reference_methods  # noqa: E402

TRAINER = pathlib.Path("/root/reference/xfmr_rec/trainer.py")
OUT = pathlib.Path(__file__).parent


def reference_compute_losses():
    tree = ast.parse(TRAINER.read_text())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "RecommenderLightningModule")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "compute_losses")
    fn.returns = None
    for a in fn.args.args:
        a.annotation = None
    ns = {"torch": torch, "loss_classes": ref_losses}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), str(TRAINER), "exec"), ns)
    return ns["compute_losses"]


def make_case(name, cfg_kwargs, n_items=300, batch=6, seq_len=12, dim=32, seed=0):
    g = torch.Generator().manual_seed(seed)
    weights = torch.cat([torch.zeros(1, dim), torch.randn((n_items, dim), generator=g) / dim ** 0.5])
    lens = torch.randint(2, seq_len + 1, (batch,), generator=g)
    valid = torch.arange(seq_len)[None, :] < lens[:, None]
    hist = torch.randint(1, n_items + 1, (batch, seq_len), generator=g) * valid
    pos = torch.randint(1, n_items + 1, (batch, seq_len), generator=g) * valid
    pos = pos * (torch.rand((batch, seq_len), generator=g) > 0.15)
    neg = torch.randint(1, n_items + 1, (batch, seq_len), generator=g) * valid
    neg[0, 1] = pos[0, 0] if pos[0, 0] != 0 else neg[0, 1]      # an in-batch negative equal to a positive
    tokens = (torch.randn((batch, seq_len, dim), generator=g) / dim ** 0.5).requires_grad_(True)
    fwd, compute_embeds = reference_methods()

    class Model(types.SimpleNamespace):
        def __call__(self, item_idx=None, *, item_embeds=None):
            return fwd(self, item_idx, item_embeds=item_embeds)

        def compute_embeds(self, h, p, n):
            return compute_embeds(self, h, p, n)

    model = Model(embeddings=torch.nn.Embedding.from_pretrained(weights, freeze=True, padding_idx=0),
                  model=StubEncoder(tokens), max_seq_length=32, device=torch.device("cpu"),
                  config=types.SimpleNamespace(is_normalized=False))
    cfg = ref_losses.LossConfig(**cfg_kwargs)
    me = types.SimpleNamespace(model=model, config=cfg,
                               loss_fns=[cls(cfg) for cls in ref_losses.LOSS_CLASSES])     # trainer.py:163-170
    out = reference_compute_losses()(me, {"history_item_idx": hist, "pos_item_idx": pos, "neg_item_idx": neg})
    out["loss/InfoNCELoss"].backward()                                                      # trainer.py:291
    logged = {k: (float(v.detach()) if isinstance(v, torch.Tensor) else float(v)) for k, v in out.items()}
    np.savez_compressed(OUT / f"compute_losses_{name}.npz", table=weights.numpy(),
                        tokens=tokens.detach().numpy(), history_item_idx=hist.numpy(),
                        pos_item_idx=pos.numpy(), neg_item_idx=neg.numpy(), cfg=json.dumps(cfg_kwargs),
                        logged=json.dumps(logged), keys=json.dumps(list(out.keys())),
                        dtokens=tokens.grad.numpy())
    print(name, len(logged), {k: round(v, 5) for k, v in list(logged.items())[:4]})


if __name__ == "__main__":
    make_case("default", {})
    make_case("scale_margin_nomask", {"scale": 8.0, "margin": 0.2, "mask_false_negatives": False}, seed=1)
