"""The scoring-and-loss train step as ONE sync-free, CUDA-graph-replayed call.

``RecommenderLightningModule.training_step`` (xfmr_rec/trainer.py:288-300) runs, after the
encoder, ``compute_embeds`` (models.py:366-419), the training loss (losses.py:128-155) and
``backward``.  Through the drop-in modules of this package that is ~10 library calls and one
device->host copy of the row counts (a boolean-mask index needs its output shape on the host;
the reference syncs the same way).  :class:`PoolLossStep` runs the same kernels through
``xr_pool_step`` instead: the counts stay on the device, every buffer is sized by ``B*L``, the
launch sequence is captured once into a CUDA graph and replayed per step.  Results are
bit-identical to ``compute_embeds`` + loss + ``backward`` on bf16 operands
(tests/test_gpu_step.py).

    step = PoolLossStep(item_embeddings, InfoNCELoss(config), batch_size=128, seq_len=200)
    loss, dtok = step(token_embeddings, history_item_idx, pos_item_idx, neg_item_idx)
    # loss: 0-dim fp32 tensor (sum over rows); dtok: dL/d token_embeddings, (B, L, D)

Inputs may live in pinned host memory: they are copied into the step's static buffers on its copy
stream, so consecutive steps of two alternating :class:`PoolLossStep` objects overlap the next
batch's host->device copy with the current batch's kernels (bench.py ``e2e``).
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _native as N
from . import ops
from .losses import EmbedLoss, _autocast_bf16
from .models import ItemEmbeddings

_STEP_KINDS = ("InfoNCELoss", "NCELoss", "PairwiseHingeLoss", "PairwiseLogisticLoss",
               "ContrastiveLoss", "AlignmentContrastiveLoss")
_COSINE_KINDS = ("ContrastiveLoss", "AlignmentContrastiveLoss")


class PoolLossStep:
    def __init__(self, embeddings: ItemEmbeddings, loss: EmbedLoss, batch_size: int, seq_len: int, *,
                 token_dtype: torch.dtype = torch.bfloat16, grad_dtype: torch.dtype | None = None,
                 want_grad: bool = True, logits_bf16: bool | None = None, use_graph: bool = True,
                 check_indices: bool = False, monitor: bool = False, pipelined: bool = False,
                 host_tokens_in_place: bool = False, monitor_one_pass: bool = True) -> None:
        name = type(loss).__name__
        if name not in _STEP_KINDS:
            raise NotImplementedError(
                f"PoolLossStep serves {_STEP_KINDS}; {name} goes through compute_embeds + the loss module")
        cfg = loss.config
        if cfg.target_position != "first":
            raise NotImplementedError("PoolLossStep: candidates are [positive | shared pool] "
                                      "(target_position='first', models.py:410)")
        if cfg.num_hard_negatives:
            raise NotImplementedError("PoolLossStep: hard-negative mining needs the materialised path")
        if cfg.scale <= 0 and (name == "InfoNCELoss" or monitor):
            raise NotImplementedError("PoolLossStep: the fused softmax needs scale > 0 "
                                      "(other scales go through compute_embeds + the loss module)")
        dev = embeddings.weight.device
        if dev.type != "cuda":
            raise N.NativeError("PoolLossStep needs the item table on a CUDA device; there is no CPU fallback")
        if not ops.fused_pool_supported(torch.empty((1, 384), dtype=torch.bfloat16, device=dev), None):
            raise N.NativeError("PoolLossStep needs the tcgen05 kernels (sm_100, D = 384)")
        assert token_dtype in (torch.float32, torch.bfloat16)
        self.device, self.b, self.l = dev, int(batch_size), int(seq_len)
        self.d = embeddings.embedding_dim
        self.kind = N.LOSS_KIND[name]
        # dot logits are bf16 under bf16-mixed autocast or for bf16 encoder output (losses.py:195)
        self.cosine = name in _COSINE_KINDS      # CCL family: row-normalised operands, fp32 logits
        if self.cosine:
            logits_bf16 = False
        if logits_bf16 is None:
            logits_bf16 = token_dtype == torch.bfloat16 or _autocast_bf16()
        self.cfg = ops.make_cfg(cfg, logits_bf16=bool(logits_bf16))
        self.table = embeddings.weight_bf16()
        self.rownz = embeddings.rownz()
        self.n_table_rows = embeddings.num_embeddings
        n = self.b * self.l
        self.n_pos = n
        grad_dtype = grad_dtype or token_dtype
        with torch.cuda.device(dev):
            self.hist = torch.zeros(n, dtype=torch.int64, device=dev)
            self.pos = torch.zeros(n, dtype=torch.int64, device=dev)
            self.neg = torch.zeros(n, dtype=torch.int64, device=dev)
            self.tok = torch.zeros((n, self.d), dtype=token_dtype, device=dev)
            self.dtok = torch.zeros((n, self.d), dtype=grad_dtype, device=dev) if want_grad else None
            self.loss_buf = torch.zeros(2, dtype=torch.float64, device=dev)
            self.counts = torch.zeros(2, dtype=torch.int64, device=dev)
            self.err = torch.zeros(1, dtype=torch.int32, device=dev) if check_indices else None
            nbytes = (N.lib().xr_pool_step_monitor_workspace_bytes(n, self.d) if (monitor or self.cosine)
                      else N.lib().xr_pool_step_workspace_bytes(n, self.d))
            # monitor: LogitsStatistics + all seven losses (trainer.py:250-263) in the same sequence
            self.monitor = bool(monitor)
            # one tensor-core pass for the train loss, its gradient AND the monitoring sums of both logit
            # families (xr_pool_step_compute_mon) instead of three; the dot-family train losses
            self.monitor_one_pass = bool(monitor and monitor_one_pass and cfg.scale > 0 and name in (
                "InfoNCELoss", "NCELoss", "PairwiseHingeLoss", "PairwiseLogisticLoss"))
            self.mon_dot = torch.zeros(N.XR_NUM_LOSSES, dtype=torch.float64, device=dev)
            self.mon_cos = torch.zeros(N.XR_NUM_LOSSES, dtype=torch.float64, device=dev)
            self.mon_stats = torch.zeros(N.XR_STATS_SLOTS, dtype=torch.float64, device=dev)
            self.ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
            off = (-self.ws.data_ptr()) % 256
            self._ws_ptr, self._ws_bytes = self.ws.data_ptr() + off, nbytes
            self.copy_stream = torch.cuda.Stream(device=dev)
            self._copied = torch.cuda.Event()
            self._done = torch.cuda.Event()
            self._done.record()
        # pipelined: load() also runs the INGEST phase (compaction, plan, gathers) on the copy stream,
        # run() only the COMPUTE phase -- with two alternating step objects the ingest of batch i+1
        # overlaps the tensor-core kernel of batch i.  Bit-identical to the one-call step.  Measured on
        # the B200 at BASELINE configs[1] it does NOT pay (0.47 against 0.39 ms per end-to-end step:
        # the ingest kernels delay CTAs of the persistent tensor-core kernel), so it is off by default;
        # it is the structure to use when the ingest side is heavier (fp32 tokens, longer sequences).  host_tokens_in_place: the gather reads pinned
        # HOST token embeddings directly (only the selected rows cross the host link) instead of a DMA
        # copy of the whole (B, L, D) tensor -- measured SLOWER on the B200 box (0.28 ms against
        # 0.09 ms for the 20 MB copy: a kernel's reads of host memory are latency-bound), so it is off.
        self.pipelined = bool(pipelined)
        self.host_tokens_in_place = bool(host_tokens_in_place)
        self._tok_src = None          # pinned host tensor the gather reads in place (kept alive)
        self._ingested = False
        self.graph = None
        self.graph_compute = None
        if use_graph:
            self._capture()

    # ------------------------------------------------------------------------------------------
    def _launch_compute_mon(self) -> None:
        N.call("xr_pool_step_compute_mon", self.n_pos, self.d, self.kind, C.byref(self.cfg), 1.0,
               ops._p(self.dtok), ops._DT[self.dtok.dtype] if self.dtok is not None else N.XR_F32,
               ops._p(self.loss_buf), ops._p(self.mon_dot), ops._p(self.mon_cos), ops._p(self.mon_stats),
               C.c_void_p(self._ws_ptr), self._ws_bytes, ops._stream())

    def _launch(self) -> None:
        if self.monitor_one_pass:
            self._launch_ingest()
            self._launch_compute_mon()
            return
        N.call("xr_pool_step", ops._p(self.hist), ops._p(self.pos), ops._p(self.neg), self.n_pos,
               ops._p(self.tok), ops._DT[self.tok.dtype], ops._p(self.table), ops._p(self.rownz),
               self.n_table_rows, self.d, self.kind, C.byref(self.cfg), 1.0, ops._p(self.dtok),
               ops._DT[self.dtok.dtype] if self.dtok is not None else N.XR_F32, ops._p(self.loss_buf),
               ops._p(self.counts), ops._p(self.err), C.c_void_p(self._ws_ptr), self._ws_bytes,
               ops._stream())
        if self.monitor:
            N.call("xr_pool_step_monitor", self.n_pos, self.d, C.byref(self.cfg), ops._p(self.mon_dot),
                   ops._p(self.mon_cos), ops._p(self.mon_stats), C.c_void_p(self._ws_ptr), self._ws_bytes,
                   ops._stream())

    def _launch_ingest(self, tok_ptr=None, tok_dtype=None) -> None:
        N.call("xr_pool_step_ingest", ops._p(self.hist), ops._p(self.pos), ops._p(self.neg), self.n_pos,
               C.c_void_p(tok_ptr) if tok_ptr is not None else ops._p(self.tok),
               ops._DT[tok_dtype if tok_dtype is not None else self.tok.dtype], ops._p(self.table),
               ops._p(self.rownz), self.n_table_rows, self.d, ops._p(self.counts), ops._p(self.err),
               C.c_void_p(self._ws_ptr), self._ws_bytes, ops._stream())

    def _launch_compute(self) -> None:
        if self.monitor_one_pass:
            self._launch_compute_mon()
            return
        N.call("xr_pool_step_compute", self.n_pos, self.d, self.kind, C.byref(self.cfg), 1.0,
               ops._p(self.dtok), ops._DT[self.dtok.dtype] if self.dtok is not None else N.XR_F32,
               ops._p(self.loss_buf), C.c_void_p(self._ws_ptr), self._ws_bytes, ops._stream())
        if self.monitor:
            N.call("xr_pool_step_monitor", self.n_pos, self.d, C.byref(self.cfg), ops._p(self.mon_dot),
                   ops._p(self.mon_cos), ops._p(self.mon_stats), C.c_void_p(self._ws_ptr), self._ws_bytes,
                   ops._stream())

    def _capture(self) -> None:
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):   # warm-up outside capture: one-time kernel attributes
                self._launch()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._launch()
            self.graph = g
            if self.pipelined:   # the compute phase alone (the ingest is issued by load())
                gc_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gc_):
                    self._launch_compute()
                self.graph_compute = gc_

    # ------------------------------------------------------------------------------------------
    def load(self, token_embeddings, history_item_idx, pos_item_idx, neg_item_idx) -> None:
        """Copy one SeqBatch (host pinned or device tensors) into the static buffers on the copy
        stream.  Waits for the previous run of THIS step object to finish with the buffers."""
        n = self.n_pos
        inputs = (token_embeddings, history_item_idx, pos_item_idx, neg_item_idx)
        dev_inputs = [t for t in inputs if t.is_cuda]
        with torch.cuda.device(self.device):
            producer = torch.cuda.current_stream()
        with torch.cuda.device(self.device), torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self._done)
            if dev_inputs:
                # device-resident inputs (the encoder output of THIS step) are still being written on
                # the caller's stream: order the copies after it, and keep the caching allocator from
                # recycling the blocks while the copy stream reads them
                self.copy_stream.wait_stream(producer)
                for t in dev_inputs:
                    t.record_stream(self.copy_stream)
            self.hist.copy_(history_item_idx.reshape(n), non_blocking=True)
            self.pos.copy_(pos_item_idx.reshape(n), non_blocking=True)
            self.neg.copy_(neg_item_idx.reshape(n), non_blocking=True)
            tok2d = token_embeddings.reshape(n, self.d)
            in_place = (self.pipelined and self.host_tokens_in_place and not tok2d.is_cuda
                        and tok2d.is_pinned() and tok2d.is_contiguous()
                        and tok2d.dtype in (torch.float32, torch.bfloat16) and tok2d.data_ptr() % 16 == 0)
            if in_place:
                self._tok_src = tok2d     # read in place by the gather kernel (kept alive until then)
            else:
                self._tok_src = None
                self.tok.copy_(tok2d, non_blocking=True)
            if self.pipelined:
                if in_place:
                    self._launch_ingest(tok2d.data_ptr(), tok2d.dtype)
                else:
                    self._launch_ingest()
                self._ingested = True
            self._copied.record()

    def run(self):
        """Enqueue the step on the current stream (after the pending load).  Returns
        (loss: 0-dim fp32 tensor, dtok: (B, L, D) | None) — views of the static buffers, valid until
        the next load()."""
        with torch.cuda.device(self.device):
            torch.cuda.current_stream().wait_event(self._copied)
            if self.pipelined and self._ingested:     # load() already ran the ingest phase
                if self.graph_compute is not None:
                    self.graph_compute.replay()
                else:
                    self._launch_compute()
                self._ingested = False
            elif self.graph is not None:
                self.graph.replay()
            else:
                self._launch()
            self._done.record()
        loss = self.loss_buf.view(torch.float32)[2]
        dtok = self.dtok.view(self.b, self.l, self.d) if self.dtok is not None else None
        return loss, dtok

    def __call__(self, token_embeddings, history_item_idx, pos_item_idx, neg_item_idx):
        self.load(token_embeddings, history_item_idx, pos_item_idx, neg_item_idx)
        return self.run()

    def enqueue(self, token_embeddings, history_item_idx, pos_item_idx, neg_item_idx):
        """``load()`` + ``run()`` for DEVICE inputs entirely on the current stream: no copy stream, no events,
        not the step's own graph.  This is the form an enclosing CUDA-graph capture can record (see
        :class:`~xfmr_rec_b200.encoder.GraphedEncoderStep`); eager use is equivalent to ``__call__``."""
        n = self.n_pos
        for t in (token_embeddings, history_item_idx, pos_item_idx, neg_item_idx):
            if not t.is_cuda:
                raise N.NativeError("PoolLossStep.enqueue takes device tensors (use __call__ for host batches)")
        with torch.cuda.device(self.device):
            self.hist.copy_(history_item_idx.reshape(n))
            self.pos.copy_(pos_item_idx.reshape(n))
            self.neg.copy_(neg_item_idx.reshape(n))
            self.tok.copy_(token_embeddings.reshape(n, self.d))
            self._tok_src = None
            self._launch()
        loss = self.loss_buf.view(torch.float32)[2]
        dtok = self.dtok.view(self.b, self.l, self.d) if self.dtok is not None else None
        return loss, dtok

    def loss_dict(self) -> tuple[dict[str, torch.Tensor], dict[str, float]]:
        """What ``compute_losses`` logs (trainer.py:250-263) for the last ``run()`` of a
        ``monitor=True`` step: ({"loss/<Name>": 0-dim fp32 tensor} for all seven losses, the
        LogitsStatistics dict).  The tensors are device-side; building the statistics dict is the
        only device->host copy (the reference does nine ``.item()`` syncs plus eight logit passes)."""
        if not self.monitor:
            raise RuntimeError("PoolLossStep(monitor=True) is needed for loss_dict()")
        from .losses import LOSS_CLASSES, stats_dict

        out = {}
        for cls in LOSS_CLASSES:
            src = self.mon_cos if cls.COSINE else self.mon_dot
            out[f"loss/{cls.__name__}"] = src[N.LOSS_KIND[cls.__name__]].to(torch.float32)
        return out, stats_dict(self.mon_stats.tolist())

    def row_counts(self) -> tuple[int, int]:
        """(M_a, M) of the last step — a device->host copy; not needed by the step itself."""
        m_a, m = self.counts.tolist()
        return int(m_a), int(m)

    def check(self) -> None:
        if self.err is not None and int(self.err.item()):
            raise IndexError("item index out of range in PoolLossStep")
