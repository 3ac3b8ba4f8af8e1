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


class ShardedIndex:
    """Catalog sharded across the ranks of a process group."""

    def __init__(self, local_index, *, group=None, merge_fn: Callable | None = None):
        self.local = local_index
        self.group = group
        self.merge_fn = merge_fn

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

    def search_batch(self, queries: torch.Tensor, exclude_rows=None, top_k: int = 20):
        s, i = self.local.search_batch(queries, exclude_rows, top_k)
        return all_gather_merge(s, i, top_k, group=self.group, merge_fn=self.merge_fn)


def reduce_loss(loss: torch.Tensor, *, group=None) -> torch.Tensor:
    """Global summed loss for logging: the reference's losses are sums over rows
    (losses.py:146-147), so ranks add."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        loss = loss.detach().clone()
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
    return loss
