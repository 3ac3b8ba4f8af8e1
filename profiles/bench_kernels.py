"""Micro-benchmarks of the HBM-bound kernel families against the measured copy bandwidth.
    python profiles/bench_kernels.py [--json out.json]
gather (family 1): algorithmic bytes = rows*D*b read + rows*D*b written (+ 8 B index per row)
top-k  (family 3): U*N*4 read + U*k*12 written
Inputs are larger than the 126 MB L2, CUDA events on the launching stream, 3 warm-ups."""
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "transformer-recommenders_b200")]
import torch

import xfmr_rec_b200 as xr

peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
HBM = peaks.get("hbm_gbs", 6650.0)
dev = torch.device("cuda", 0)


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = []
g = torch.Generator(device=dev).manual_seed(0)
# ---- gather: ML-32M-shaped table (87,586 x 384), 1M gathered rows (768 MB bf16 out) -------------
for dt, b in ((torch.bfloat16, 2), (torch.float32, 4)):
    table = torch.randn(87586, 384, device=dev, generator=g).to(dt)
    rows = 1_000_000 if dt == torch.bfloat16 else 500_000
    idx = torch.randint(0, 87586, (rows,), device=dev, generator=g)
    ms = timeit(lambda: xr.ops.gather_rows(table, idx))
    byts = rows * 384 * b * 2 + rows * 8
    out.append({"kernel": f"gather_rows {dt}".replace("torch.", ""), "rows": rows, "ms": ms,
                "GB/s": byts / ms / 1e6, "frac_of_measured_hbm": byts / ms / 1e6 / HBM,
                "note": "table (67/134 MB) is L2-resident: read side is L2 traffic, write side HBM"})
# catalog-sized table: reads also come from HBM
table = torch.randn(4_000_000, 384, device=dev, generator=g).bfloat16()
idx = torch.randint(0, 4_000_000, (1_000_000,), device=dev, generator=g)
ms = timeit(lambda: xr.ops.gather_rows(table, idx))
byts = 1_000_000 * 384 * 2 * 2 + 1_000_000 * 8
out.append({"kernel": "gather_rows bfloat16 (3 GB table)", "rows": 1_000_000, "ms": ms,
            "GB/s": byts / ms / 1e6, "frac_of_measured_hbm": byts / ms / 1e6 / HBM})
del table
# ---- top-k over a materialised score matrix -------------------------------------------------------
for u, n in ((1, 10_000_000), (64, 10_000_000), (256, 2_000_000), (4096, 262_144)):
    s = torch.randn(u, n, device=dev, generator=g)
    ms = timeit(lambda: xr.ops.topk(s, 100), reps=5)
    byts = u * n * 4 + u * 100 * 12
    out.append({"kernel": "topk k=100", "U": u, "N": n, "ms": ms, "GB/s": byts / ms / 1e6,
                "frac_of_measured_hbm": byts / ms / 1e6 / HBM})
    del s
for r in out:
    print(json.dumps(r))
if "--json" in sys.argv:
    pathlib.Path(sys.argv[sys.argv.index("--json") + 1]).write_text(json.dumps(out, indent=1))
