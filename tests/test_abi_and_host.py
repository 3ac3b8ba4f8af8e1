"""CPU tests: the C-ABI library loads and exports every declared symbol; host-side logic."""

import pathlib
import re

import numpy as np
import pytest
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    from xfmr_rec_b200 import _native

    header = (ROOT / "include" / "xfmr_b200.h").read_text()
    declared = set(re.findall(r"\b(xr_[a-z0-9_]+)\s*\(", header))
    declared -= {"xr_loss_config"}
    assert declared, "no declarations parsed"
    lib = _native.lib()               # raises if the .so is missing or lacks a bound symbol
    for name in declared:
        assert hasattr(lib, name), name
    assert declared <= set(_native.PROTOTYPES), declared - set(_native.PROTOTYPES)
    assert lib.xr_abi_version() == 2


def test_no_cpu_fallback():
    import xfmr_rec_b200 as xr

    q, cand = torch.randn(4, 8), torch.randn(4, 5, 8)
    with pytest.raises(xr._native.NativeError):
        xr.InfoNCELoss(xr.LossConfig())(q, cand)
    with pytest.raises(xr._native.NativeError):
        xr.ops.gather_rows(torch.randn(4, 8), torch.zeros(2, dtype=torch.int64))
    with pytest.raises(xr._native.NativeError):
        xr.ops.topk(torch.randn(2, 10), 3)


def test_product_never_imports_oracle():
    pkg = ROOT / "transformer-recommenders_b200" / "xfmr_rec_b200"
    for f in pkg.glob("*.py"):
        src = f.read_text()
        assert "oracle" not in src.replace("stable-sort oracle", ""), f.name


def test_registry_and_config_mirror_reference():
    import xfmr_rec_b200 as xr

    assert [c.__name__ for c in xr.LOSS_CLASSES] == [
        "AlignmentLoss", "AlignmentContrastiveLoss", "ContrastiveLoss", "InfoNCELoss", "NCELoss",
        "PairwiseHingeLoss", "PairwiseLogisticLoss"]
    cfg = xr.LossConfig()
    assert (cfg.target_position, cfg.mask_false_negatives, cfg.num_hard_negatives, cfg.scale,
            cfg.margin) == ("first", True, 0, 1.0, 0.5)
    for cls in xr.LOSS_CLASSES:
        mod = cls(cfg)
        assert isinstance(mod, torch.nn.Module) and not list(mod.parameters()) and not list(mod.buffers())
    # any object with the five attributes is accepted (trainer passes its LightningConfig)
    class Cfg:
        target_position, mask_false_negatives, num_hard_negatives, scale, margin = "first", True, 0, 1.0, 0.5
    xr.InfoNCELoss(Cfg())


def test_check_embeds_and_target_assertions_on_host():
    import xfmr_rec_b200 as xr

    loss = xr.InfoNCELoss(xr.LossConfig())
    with pytest.raises(AssertionError):
        loss.check_embeds(torch.zeros(3), torch.zeros(3, 2, 4))
    with pytest.raises(AssertionError):
        loss.check_embeds(torch.zeros(3, 4), torch.zeros(3, 4))
    with pytest.raises(AssertionError):
        loss.check_embeds(torch.zeros(3, 4), torch.zeros(2, 2, 4))
    with pytest.raises(AssertionError):
        loss.check_embeds(torch.zeros(3, 4), torch.zeros(3, 2, 5))
    with pytest.raises(AssertionError):
        loss.check_target(3, 2, torch.zeros(3, dtype=torch.long))
    with pytest.raises(AssertionError):
        xr.InfoNCELoss(xr.LossConfig(target_position=None)).check_target(3, 2, None)
    with pytest.raises(AssertionError):
        xr.InfoNCELoss(xr.LossConfig(target_position=None)).check_target(3, 2, torch.zeros(2, dtype=torch.long))
    h = xr.PoolCandidates(torch.zeros(5, 8), torch.zeros(9, 8))
    assert h.dim() == 3 and tuple(h.size()) == (5, 10, 8) and h.size(1) == 10
    assert tuple(h[torch.tensor([True, False, True, True, False])].size()) == (3, 10, 8)
    assert tuple(h.dense().shape) == (5, 10, 8)


def test_shard_ranges_cover_catalog():
    from xfmr_rec_b200.dist import shard_range

    for n in (1, 127, 128, 1000, 10_000_000):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b


def _worker(rank, world, port, out):
    import torch.distributed as dist

    from oracle import xfmr_oracle as orc
    from xfmr_rec_b200.dist import all_gather_merge, reduce_loss, shard_range

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    u, n, k = 5, 1000, 20
    s = np.round(rng.standard_normal((u, n)).astype(np.float32) * 4) / 4   # identical on all ranks
    lo, hi = shard_range(n, rank, world)
    ls, li = orc.topk_rows(s[:, lo:hi], k)                                  # local step (injected)

    def merge(cs, ci, kk):                                                  # merge step (injected)
        cs, ci = cs.numpy(), ci.numpy()
        order = np.lexsort((ci, -cs), axis=1)[:, :kk]
        return (torch.from_numpy(np.take_along_axis(cs, order, 1)),
                torch.from_numpy(np.take_along_axis(ci, order, 1)))

    ms, mi = all_gather_merge(torch.from_numpy(ls), torch.from_numpy(li + lo), k, merge_fn=merge)
    ws, wi = orc.topk_rows(s, k)
    ok = np.array_equal(mi.numpy(), wi) and np.array_equal(ms.numpy(), ws)
    tot = reduce_loss(torch.tensor(float(rank + 1)))
    ok = ok and float(tot) == sum(range(1, world + 1))
    out[rank] = ok
    dist.destroy_process_group()


def test_sharded_merge_world2_gloo():
    """N>1 host logic on CPU: shard ranges + all-gather + merge order == single-rank top-k."""
    import torch.multiprocessing as mp

    world, port = 2, 29731
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the driver's reference arm): one JSON line with the base contract's
    keys, `impl: reference`, a cpu_baseline describing the run and a zero-copy e2e block; runs on
    the host cores only."""
    import json
    import pathlib
    import subprocess
    import sys

    root = pathlib.Path(__file__).resolve().parent.parent
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--gpus", "1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "seq/s"
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, str(root))
    import bench

    assert d["config"] == bench.config_dict(1) and d["metric"] == bench.METRIC   # same config in both arms


def test_torch_library_ops_registered_with_fake_impls():
    """the entry points are dispatcher-visible ``xfmr_b200::*`` ops registered at import,
    each with a fake (meta) implementation; CUDA is the only backend (no CPU fallback)."""
    from torch._subclasses.fake_tensor import FakeTensorMode

    import xfmr_rec_b200 as xr

    for name in xr.ops.CUSTOM_OPS:
        assert hasattr(torch.ops.xfmr_b200, name), name
    with FakeTensorMode():
        q = torch.empty((10, 384), dtype=torch.bfloat16, device="cuda")
        neg = torch.empty((33, 384), dtype=torch.bfloat16, device="cuda")
        loss, dq = torch.ops.xfmr_b200.pool_loss(q, q, neg, 3, True, 1.0, 0.5, True)
        assert loss.shape == () and dq.shape == (10, 384) and dq.dtype == torch.float32
        s, i = torch.ops.xfmr_b200.score_topk(q, neg, 5)
        assert s.shape == (10, 5) and i.dtype == torch.int64
        g = torch.ops.xfmr_b200.gather_rows(neg, torch.empty((4, 7), dtype=torch.int64, device="cuda"))
        assert g.shape == (4, 7, 384)
        m, v = torch.ops.xfmr_b200.retrieval_metrics(torch.empty((6, 20), dtype=torch.int64, device="cuda"),
                                                     torch.empty(7, dtype=torch.int64, device="cuda"),
                                                     torch.empty(9, dtype=torch.int64, device="cuda"), 20)
        assert m.shape == (6, 7) and v.shape == (6,)
    with pytest.raises(NotImplementedError):
        torch.ops.xfmr_b200.topk(torch.zeros(2, 5), 2)


def test_service_wire_types_round_trip():
    """service.py:30-72: same field names / defaults, arrays travel as float lists."""
    from xfmr_rec_b200 import service as S

    q = S.Query(embedding=[0.1, 0.2], exclude_item_ids=["3"], top_k=5)
    assert S.Query().top_k == 20 and S.Query().embedding is None
    back = S.Query.model_validate_json(q.model_dump_json())
    assert isinstance(back.embedding, np.ndarray) and back.embedding.dtype == np.float32
    assert back.exclude_item_ids == ["3"] and back.top_k == 5
    c = S.ItemCandidate(item_id="1", item_text="x", score=0.5)
    assert c.model_dump() == {"item_id": "1", "item_text": "x", "score": 0.5}
    u = S.UserQuery(history=S.Activity(item_id=["1"], item_text=["a"]))
    assert u.user_id == "0" and u.history.item_id == ["1"] and u.target is None
    i = S.ItemQuery.model_validate({"item_id": "7", "item_text": "t", "embedding": np.ones(3)})
    assert i.model_dump()["embedding"] == [1.0, 1.0, 1.0]


def test_synthetic_batch_generators_agree():
    """bench.py draws its inputs from the package's generator; the parity tests use the oracle's: same
    algorithm, same arrays."""
    from oracle import xfmr_oracle as orc
    from xfmr_rec_b200.data import synthetic_batch

    a, b = orc.synth_batch(300, 5, 17, dim=32, seed=4), synthetic_batch(300, 5, 17, dim=32, seed=4)
    assert a.keys() == b.keys()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
