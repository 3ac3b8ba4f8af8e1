"""Multi-GPU retrieval exchange (needs >= 2 GPUs on the box; skipped otherwise): the NVLink
peer-memory merge and the NCCL all-gather path both reproduce the unsharded search bit for bit."""

import json
import pathlib
import subprocess
import sys

import pytest
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.mark.gpu
def test_peer_memory_and_nccl_exchange_equal_unsharded_search():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n_gpus = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_gpus}",
           "--master-addr", "127.0.0.1", "--master-port", "29541",
           str(ROOT / "profiles" / "run_peer_exchange.py"), "300000", "130"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["nccl"]["equals_unsharded"] and res["peer"]["equals_unsharded"]
