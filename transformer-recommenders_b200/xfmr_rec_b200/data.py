"""SeqBatch construction on the device — mirror of ``SeqDataset`` (xfmr_rec/data.py:558-805).

The reference builds every training example in Python (``__getitem__``, data.py:749-785):
``sample_sequence`` (:669-689), ``sample_positives`` (:691-721, a Python loop per position) and
``sample_negatives`` (:723-747, a set difference over the WHOLE catalog per example), then pads
them in ``collate`` (:787-805).  Here the per-user event histories live on the GPU in CSR form and
one kernel launch (``xr_seq_sample_batch``) produces the three ``(B, max_seq_length)`` index
tensors of a ``SeqBatch`` (data.py:534-540) — already on the device, ready for
``compute_embeds`` / ``PoolLossStep``, no host→device copy of indices.

Same configuration names as ``SeqDataConfig`` (data.py:543-545).  Randomness is counter-based
(Philox4x32-10 keyed by ``(seed, step, row)``), so a batch is reproducible from three integers and
independent of which other rows share the batch.  There is no CPU path.
"""

from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np
import torch

from . import _native as N
from .ops import _on, _p, _require_cuda, _stream


@dataclasses.dataclass
class SeqDataConfig:
    """xfmr_rec/data.py:543-545."""

    max_seq_length: int = 32
    pos_lookahead: int = 0


def build_history_csr(histories, labels, num_items: int, max_seq_length: int) -> dict:
    """Host-side layout of the per-user histories for ``xr_seq_sample_batch`` (pure numpy):
    empty histories dropped (data.py:652), CSR offsets / items / labels, the inclusive count of
    positive labels inside each history, the sorted unique items of each history, and the dataset-row
    -> history map of ``duplicate_rows`` (data.py:618-636: ``(len - 1) // max_seq_length + 1`` rows per
    history)."""
    keep = [i for i, h in enumerate(histories) if len(h) > 0]
    hs = [np.asarray(histories[i], np.int64) for i in keep]
    ls = [np.asarray(labels[i], bool) for i in keep]
    for h, l in zip(hs, ls):
        if len(h) != len(l):
            raise ValueError("history and label arrays differ in length")
        if h.min() < 1 or h.max() > num_items:
            raise IndexError("history item index out of range [1, num_items]")
    lens = np.array([len(h) for h in hs], np.int64)
    off = np.zeros(len(hs) + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    uniq = [np.unique(h) for h in hs]
    uoff = np.zeros(len(hs) + 1, np.int64)
    np.cumsum([len(u) for u in uniq], out=uoff[1:])
    copies = (lens - 1) // int(max_seq_length) + 1
    return {
        "kept": keep,
        "hist_off": off,
        "items": np.concatenate(hs) if hs else np.zeros(0, np.int64),
        "labels": np.concatenate(ls).astype(np.uint8) if ls else np.zeros(0, np.uint8),
        "pos_prefix": np.concatenate([np.cumsum(l, dtype=np.int32) for l in ls]) if ls else np.zeros(0, np.int32),
        "uniq_off": uoff,
        "uniq_items": np.concatenate(uniq) if uniq else np.zeros(0, np.int64),
        "row_hist": np.repeat(np.arange(len(hs), dtype=np.int64), copies),
    }


class SeqBatchSampler:
    """Device-resident stand-in for ``SeqDataset`` + ``collate``.

    ``histories[u]`` / ``labels[u]`` are one user's ``history_item_idx`` (1-based item indices,
    0 is padding and never appears) and ``history_label`` AFTER ``map_id2idx`` (data.py:589-616:
    unknown items dropped, events after the last positive trimmed).  Empty histories are dropped
    (data.py:652) and long ones are repeated ``(len - 1) // max_seq_length + 1`` times
    (``duplicate_rows``, data.py:618-636) — as a row→history map, not as copies.
    """

    def __init__(self, config, histories, labels, num_items: int, device="cuda", seed: int = 0):
        self.config = config
        self.num_items = int(num_items)
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise N.NativeError("SeqBatchSampler needs a CUDA device; there is no CPU fallback")
        csr = build_history_csr(histories, labels, self.num_items, int(config.max_seq_length))
        self.kept_histories = csr["kept"]
        self.num_histories = len(csr["kept"])
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)  # noqa: E731
        self.hist_off, self.items = up(csr["hist_off"]), up(csr["items"])
        self.labels, self.pos_prefix = up(csr["labels"]), up(csr["pos_prefix"])
        self.uniq_off, self.uniq_items, self.row_hist = up(csr["uniq_off"]), up(csr["uniq_items"]), up(csr["row_hist"])
        row_hist = csr["row_hist"]
        self._row_hist_host = row_hist

    def __len__(self) -> int:  # data.py:659-666
        return int(self._row_hist_host.shape[0])

    def sample(self, rows: torch.Tensor, step: int = 0, *, return_lengths: bool = False):
        """``collate([dataset[r] for r in rows])`` (data.py:749-805) -> ``SeqBatch`` index tensors
        ``(B, max_seq_length)`` int64 on the device, right-padded with 0.  (The reference pads to
        the longest sequence of the batch; the extra all-padding columns are inert: every consumer
        masks on ``idx != 0``, models.py:343, 413.)"""
        _require_cuda(rows)
        rows = rows.to(torch.int64).contiguous()
        b, L = rows.numel(), int(self.config.max_seq_length)
        dev = self.device
        out = torch.empty((3, b, L), dtype=torch.int64, device=dev)
        lens = torch.empty(b, dtype=torch.int32, device=dev)
        with _on(dev):
            N.call("xr_seq_sample_batch", _p(self.hist_off), _p(self.items), _p(self.labels),
                   _p(self.pos_prefix), _p(self.uniq_off), _p(self.uniq_items), _p(self.row_hist),
                   _p(rows), b, self.num_items, L, int(self.config.pos_lookahead),
                   C.c_uint64(self.seed), C.c_uint64(int(step) & 0xFFFFFFFFFFFFFFFF), _p(out[0]),
                   _p(out[1]), _p(out[2]), _p(lens), _stream())
        batch = {"history_item_idx": out[0], "pos_item_idx": out[1], "neg_item_idx": out[2]}
        if return_lengths:
            batch["seq_len"] = lens
        return batch

    def epoch(self, batch_size: int, epoch: int = 0, *, shuffle: bool = True, drop_last: bool = False):
        """Iterate one epoch of batches (the DataLoader of data.py:905-927 with ``shuffle=True``);
        the permutation is drawn on the device from ``(seed, epoch)``."""
        n = len(self)
        if shuffle:
            g = torch.Generator(device=self.device).manual_seed((self.seed * 1_000_003 + epoch) % (1 << 63))
            perm = torch.randperm(n, generator=g, device=self.device)
        else:
            perm = torch.arange(n, device=self.device)
        n_batches = n // batch_size if drop_last else -(-n // batch_size)
        for i in range(n_batches):
            yield self.sample(perm[i * batch_size:(i + 1) * batch_size], step=epoch * n_batches + i)


def synthetic_batch(n_items: int, batch: int, seq_len: int, dim: int = 384, seed: int = 0,
                    pos_pad_frac: float = 0.05, table=None) -> dict:
    """MovieLens-shaped synthetic SeqBatch + item table + encoder-output stand-in: table
    rows ~ N(0, 1/dim) with a zero padding row 0; sequence lengths uniform in [1, seq_len], right-padded
    with 0 as ``pad_sequence`` does (data.py:801); 5 % of the positives set to 0 (positions with no
    future positive, data.py:710-721).  numpy arrays; deterministic in ``seed``."""
    import math

    import numpy as np

    rng = np.random.default_rng(seed)
    if table is None:
        table = (rng.standard_normal((n_items + 1, dim)) / math.sqrt(dim)).astype(np.float32)
        table[0] = 0.0
    lens = rng.integers(1, seq_len + 1, size=batch)
    hist = np.zeros((batch, seq_len), np.int64)
    pos = np.zeros((batch, seq_len), np.int64)
    neg = np.zeros((batch, seq_len), np.int64)
    for b in range(batch):
        n = int(lens[b])
        hist[b, :n] = rng.integers(1, n_items + 1, size=n)
        pos[b, :n] = rng.integers(1, n_items + 1, size=n)
        neg[b, :n] = rng.integers(1, n_items + 1, size=n)
        drop = rng.random(n) < pos_pad_frac
        pos[b, :n][drop] = 0
    tokens = (rng.standard_normal((batch, seq_len, dim)) / math.sqrt(dim)).astype(np.float32)
    return {"table": table, "history_item_idx": hist, "pos_item_idx": pos,
            "neg_item_idx": neg, "token_embeddings": tokens}
