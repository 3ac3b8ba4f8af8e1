"""Multi-GPU plumbing (one process per GPU, ``torch.distributed``; NCCL on the B200 box).

The hot path shards in two places:
  * full-catalog retrieval — catalog rows are split into contiguous ranges, queries are
    replicated, every rank produces its local top-k with GLOBAL row ids, one all-gather of
    (U, k) {score fp32, id int64} follows, then a local merge under the same total order
    (score desc, id asc) — bit-identical to the single-GPU result;
  * training — data parallel over sequences, item table replicated (it is frozen); the loss
    is a SUM over rows, so the logged global loss is an all-reduce(sum); encoder gradients are
    all-reduced by the trainer's DDP wrapper, which is outside this path.
The reference has no collective call site of its own; no other exchange exists.
"""

from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous near-equal row ranges; multiples of 128 rows except for the last shard."""
    per = (n + world - 1) // world
    per = (per + 127) // 128 * 128
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def all_gather_merge(local_scores: torch.Tensor, local_ids: torch.Tensor, k: int, *, group=None,
                     merge_fn: Callable | None = None):
    """All-gather per-shard (U,k) results and merge them.  ``merge_fn(scores (U,G*k), ids, k)``
    defaults to the CUDA merge kernel (ops.topk_merge)."""
    if merge_fn is None:
        from . import ops

        merge_fn = ops.topk_merge
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return merge_fn(local_scores, local_ids, k)
    u = local_scores.size(0)
    gs = torch.empty((world * u, k), dtype=local_scores.dtype, device=local_scores.device)
    gi = torch.empty((world * u, k), dtype=local_ids.dtype, device=local_ids.device)
    dist.all_gather_into_tensor(gs, local_scores.contiguous(), group=group)
    dist.all_gather_into_tensor(gi, local_ids.contiguous(), group=group)
    cat_s = gs.view(world, u, k).permute(1, 0, 2).reshape(u, world * k)
    cat_i = gi.view(world, u, k).permute(1, 0, 2).reshape(u, world * k)
    return merge_fn(cat_s, cat_i, k)


class PeerExchange:
    """The retrieval exchange over NVLink peer memory instead of NCCL: every rank keeps its per-shard
    ``(U, k)`` lists in a symmetric-memory buffer (CUDA IPC mappings set up once by
    ``torch.distributed._symmetric_memory``), one device-side barrier follows, and each rank's merge
    kernel (``xr_topk_merge_peers``) loads all peers' lists straight through the peer mappings while
    it selects: all-gather + merge in ONE kernel, no staging copy, no permute.

    Two buffer sets alternate: search n+1's barrier on a rank is stream-ordered after its merge of
    search n, so once barrier n+1 completes everywhere every merge of search n has finished and its
    buffer set may be rewritten at search n+2 — one barrier per search is enough."""

    def __init__(self, max_queries: int, k: int, device, group=None):
        import torch.distributed._symmetric_memory as symm

        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.u_max, self.k = int(max_queries), int(k)
        self.slot_bytes = -(-(self.u_max * self.k * 12) // 256) * 256    # fp32 scores | int64 ids
        self.buf = symm.empty(2 * self.slot_bytes, dtype=torch.uint8, device=device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.turn = 0

    def gather_merge(self, local_scores: torch.Tensor, local_ids: torch.Tensor, k: int):
        import ctypes as C

        from . import _native as N, ops

        u = local_scores.size(0)
        assert u <= self.u_max and k == self.k, "PeerExchange was sized for other (U, k)"
        dev = local_scores.device
        base = self.turn * self.slot_bytes
        self.turn ^= 1
        id_off = -(-(u * k * 4) // 16) * 16
        mine_s = self.buf[base:base + u * k * 4].view(torch.float32).view(u, k)
        mine_i = self.buf[base + id_off:base + id_off + u * k * 8].view(torch.int64).view(u, k)
        mine_s.copy_(local_scores)
        mine_i.copy_(local_ids)
        self.hdl.barrier(channel=0)    # every rank's lists are in place (and search n-1's merges are done)
        sp = (C.c_void_p * self.world)(*[p + base for p in self.ptrs])
        ip = (C.c_void_p * self.world)(*[p + base + id_off for p in self.ptrs])
        out_s = torch.empty((u, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((u, k), dtype=torch.int64, device=dev)
        ws = ops._ws(N.lib().xr_topk_merge_workspace_bytes(u, k), dev)
        with ops._on(dev):
            N.call("xr_topk_merge_peers", sp, ip, self.world, u, k, ops._p(out_s), ops._p(out_i),
                   ops._p(ws), ops._stream())
        return out_s, out_i


class ShardedIndex:
    """Catalog sharded across the ranks of a process group.  ``exchange="peer"`` merges through
    NVLink peer memory (:class:`PeerExchange`), ``"nccl"`` through two all-gathers + a merge;
    ``"auto"`` tries peer memory first.  Both give the single-GPU result bit for bit."""

    def __init__(self, local_index, *, group=None, merge_fn: Callable | None = None,
                 exchange: str = "nccl", use_plan: bool = False, plan_max_exclusions: int = 0):
        self.local = local_index
        self.group = group
        self.merge_fn = merge_fn
        self.exchange = exchange
        # use_plan: the local search runs as one CUDA-graph replay (ExactIndex.compile_search), compiled
        # once per (U, top_k); exclusion lists (CSR tensor pairs of at most plan_max_exclusions ids per
        # query) pass through the plan's static buffers
        self.use_plan = use_plan
        self.plan_max_exclusions = int(plan_max_exclusions)
        self._plans: dict = {}
        self._peer: PeerExchange | None = None
        self._peer_failed: str | None = None

    def _peer_exchange(self, u: int, k: int, device):
        if self._peer is not None and (u > self._peer.u_max or k != self._peer.k):
            self._peer = None   # re-size (collective: every rank sees the same (U, k))
        if self._peer is None and self._peer_failed is None:
            try:
                self._peer = PeerExchange(max(u, 256), k, device, self.group)
            except Exception as e:  # symmetric memory not available on this system
                if self.exchange == "peer":
                    raise
                self._peer_failed = f"{type(e).__name__}: {e}"
        return self._peer

    @classmethod
    def from_catalog(cls, embeddings: torch.Tensor, config=None, device=None, *, group=None):
        """Every rank passes the same (N, D) matrix (or a factory); only its shard is kept."""
        from .index import ExactIndex

        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        lo, hi = shard_range(embeddings.size(0), rank, world)
        idx = ExactIndex(config, device, row_offset=lo)
        idx.set_catalog(embeddings[lo:hi])
        return cls(idx, group=group)

    def plans_overflowed(self) -> bool:
        """True if a graph-replayed local search issued with ``check=False`` could not be served exactly
        (see ``SearchPlan``); clears the flags."""
        return any([p.overflowed() for p in self._plans.values()])

    def search_batch(self, queries: torch.Tensor, exclude_rows=None, top_k: int = 20, *, check: bool = True):
        planned = (self.use_plan and queries.dim() == 2 and
                   (exclude_rows is None or (isinstance(exclude_rows, tuple) and self.plan_max_exclusions > 0)))
        if planned:
            mx = self.plan_max_exclusions if exclude_rows is not None else 0
            key = (queries.size(0), top_k, mx)
            if key not in self._plans:
                self._plans[key] = self.local.compile_search(queries.size(0), top_k, mx)
            s, i = self._plans[key](queries, exclude_rows, check=check)
        else:
            s, i = self.local.search_batch(queries, exclude_rows, top_k)
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1 and self.merge_fn is None:
            return s, i          # one shard: the local result is the result
        if world > 1 and self.exchange in ("peer", "auto") and self.merge_fn is None and s.is_cuda:
            px = self._peer_exchange(s.size(0), top_k, s.device)
            if px is not None:
                return px.gather_merge(s, i, top_k)
        return all_gather_merge(s, i, top_k, group=self.group, merge_fn=self.merge_fn)


def reserve_sms(n: int) -> int:
    """Leave ``n`` SMs free for concurrent kernels (the NCCL all-reduce of the trainer's DDP wrapper): the
    persistent tensor-core kernels of this package then launch ``#SMs - n`` CTAs.  Call before building
    ``PoolLossStep`` / ``SearchPlan`` objects (their CUDA graphs bake the grid).  Returns the previous value."""
    from . import _native as N

    return int(N.lib().xr_reserve_sms(int(n)))


def reduce_loss(loss: torch.Tensor, *, group=None) -> torch.Tensor:
    """Global summed loss for logging: the reference's losses are sums over rows
    (losses.py:146-147), so ranks add."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        loss = loss.detach().clone()
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
    return loss
