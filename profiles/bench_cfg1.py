"""BASELINE configs[0] (the reference's CPU-runnable case: ML-1M-shaped, 3,706 items, B=128 x L=50, InfoNCE) run
ON the GPU through the drop-in modules: fp32 (the 1e-5 exact path: fp32 CUDA-core GEMM + row losses) and bf16
(tcgen05 fused path).  Small problem: host dispatch dominates.
    python profiles/bench_cfg1.py > profiles/cfg1_r01.jsonl"""
import sys
import pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch, json
import xfmr_rec_b200 as xr
from oracle import xfmr_oracle as orc
b = orc.synth_batch(3706, 128, 50, dim=384, seed=0)
emb = xr.models.ItemEmbeddings(torch.from_numpy(b["table"]), add_padding_row=False).cuda()
tok0 = torch.from_numpy(b["token_embeddings"]).cuda()
idx = [torch.from_numpy(b[k]).cuda() for k in ("history_item_idx", "pos_item_idx", "neg_item_idx")]
fn = xr.InfoNCELoss(xr.LossConfig())
def step(dtype):
    tok = tok0.to(dtype).detach().requires_grad_(True)
    out = xr.models.compute_embeds(emb, tok, *idx, candidate_dtype=dtype)
    loss = fn(out["query_embed"], out["candidate_embed"]); loss.backward(); return loss, out
for dtype in (torch.float32, torch.bfloat16):
    for _ in range(3): l, out = step(dtype)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): step(dtype)
    e.record(); torch.cuda.synchronize()
    m, c = out["query_embed"].size(0), out["candidate_embed"].neg.size(0)
    ms = a.elapsed_time(e) / 20
    print(json.dumps({"config": "cfg1 on GPU (ML-1M-shaped, B=128 x L=50)", "dtype": str(dtype), "M": m, "C": c + 1, "step_ms": ms,
                      "seq_per_s": 128 / ms * 1e3, "GFLOP": 4 * m * (c + 1) * 384 / 1e9, "TFLOP/s": 4 * m * (c + 1) * 384 / ms / 1e9, "loss": float(l)}))
