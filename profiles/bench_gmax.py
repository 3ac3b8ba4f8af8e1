"""Retrieval scoring kernel alone (xr_score_groupmax): CTA-pair (cta_group::2) vs single-CTA variant.
    python profiles/bench_gmax.py [catalog_rows]"""
import ctypes
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch

from xfmr_rec_b200 import _native as N, ops

peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
HBM, TF = peaks.get("hbm_gbs", 6650.0), peaks.get("bf16_tflops_sustained", 1400.0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
cat = torch.randn((n, 384), generator=g, device=dev).bfloat16()
lib = N.lib()
for u in (256, 1024, 4096, 128):
    q = torch.randn((u, 384), generator=g, device=dev).bfloat16()
    for single in (0, 1):
        if u <= 128 and not single:
            continue
        lib.xr_fused_wait_stats(4 if single else 0, None)
        for _ in range(2):
            ops.score_groupmax(q, cat)
        torch.cuda.synchronize()
        lib.xr_fused_profile(1)
        for _ in range(5):
            ops.score_groupmax(q, cat)
        buf = (ctypes.c_float * 512)()
        k = lib.xr_fused_profile_read(buf, 512)
        lib.xr_fused_profile(0)
        ms = sum(buf[i] for i in range(k)) / k
        fl, by = 2.0 * u * n * 384, n * 384 * 2.0
        print(json.dumps({"kernel": "score_groupmax", "variant": "single-CTA" if single else "CTA pairs (cta_group::2)",
                          "U": u, "N": n, "ms": ms, "TFLOP/s": fl / ms / 1e9, "frac_of_measured_bf16": fl / ms / 1e9 / TF,
                          "catalog_GB/s": by / ms / 1e6, "frac_of_measured_hbm": by / ms / 1e6 / HBM}))
lib.xr_fused_wait_stats(0, None)
