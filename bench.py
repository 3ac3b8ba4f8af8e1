#!/usr/bin/env python
"""bench.py — scoring-and-loss train step (BASELINE configs[1]) on N B200s, with the other BASELINE
configurations as keyed legs of the same JSON line.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Headline workload (`config.workload`): MovieLens-20M-shaped synthetic batch — 27,278 items, 384-d frozen
item table, B=128 sequences of up to L=200 positions, InfoNCE (sampled softmax) with the in-batch shared
negative pool, bf16 tensor-core arithmetic.  One step = one pass of the hot path over one batch: index
compaction + the three embedding gathers (compute_embeds, models.py:366-419) + fused contraction / loss /
gradient (losses.py:150-155, 479-488) + the scatter of dL/dquery back to the encoder-output layout.  The
sequence encoder is outside the path (north_star); a random (B, L, 384) tensor stands in for its output.

`value`   : sequences/s with every input already resident in HBM.
`e2e`     : the same step through the public API with HOST (pinned) inputs: per step the index tensors and
            the encoder-output stand-in are copied host->device and the loss scalar is read back.
`roofline`: the fused tcgen05 kernel alone, timed per launch with CUDA events on its stream,
            4*M*C*D algorithmic FLOPs (scores + dQ) against the measured bf16 peaks.
`configs` : BASELINE configs[2] (CCL, 512 sampled negatives, 6,400 rows per GPU), configs[3] (`retrieval`:
            full-catalog top-100 with 20-200 excluded ids per query, U in {1, 256, 4096}), two configs[4]
            points (fused vs materialised logits) and the HBM-bound kernels (gather, top-k), each with its
            own roofline measured in this run.
`cpu_baseline` / `--impl reference`: the reference's arithmetic in torch CPU ops on all host threads
            (oracle/cpu_baseline.py; the reference is Python and cannot travel to the GPU box).
"""

from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import pathlib
import statistics
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
for _p in (ROOT, ROOT / "transformer-recommenders_b200"):
    if str(_p) not in sys.path:
        sys.path.insert(0, str(_p))

N_ITEMS, BATCH, SEQ_LEN, DIM = 27278, 128, 200, 384
WORKLOAD = ("ML-20M-shaped synthetic (27,278 items, 384-d), B=128 x L=200, InfoNCE in-batch "
            "shared-pool negatives, bf16 (BASELINE configs[1])")
METRIC = "train seq/sec (scoring-and-loss step)"


def config_dict(world: int) -> dict:
    """The SAME keys and values in both arms (`--impl ours` / `--impl reference`)."""
    return {"workload": WORKLOAD, "n_items": N_ITEMS, "dim": DIM, "global_batch": world * BATCH,
            "seq_len": SEQ_LEN, "loss": "InfoNCELoss", "seed": 0}


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"tf_sustained": d.get("bf16_tflops_sustained", 1400.0), "tf_burst": d.get("bf16_tflops", 1590.0),
                "hbm": d.get("hbm_gbs", 6650.0), "src": "measured (MEASURED_PEAKS.json)"}
    return {"tf_sustained": 1400.0, "tf_burst": 1590.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe): NVML polled from
    a thread every 2 ms (or an `nvidia-smi -lms` child process); samples are attributed to the timed
    window by their wall-clock arrival time."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None
        self.t0 = self.t1 = None

    def _start_nvml(self, phys: int) -> bool:
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = (("hw_slowdown", pynvml.nvmlClocksEventReasonHwSlowdown),
                    ("hw_thermal_slowdown", pynvml.nvmlClocksEventReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", pynvml.nvmlClocksEventReasonSwThermalSlowdown),
                    ("sw_power_cap", pynvml.nvmlClocksEventReasonSwPowerCap))
            pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            return False
        self._stop_flag = False

        def pump():
            while not self._stop_flag:
                try:
                    sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    row = [str(sm), str(mx)] + ["Active" if r & b else "Not Active" for _, b in bits]
                    self.rows.append((time.time(), row))
                except Exception:
                    pass
                time.sleep(self.period)

        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()
        return True

    def start(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [x.strip() for x in vis.split(",")] if vis else []
        phys = ids[self.gpu] if self.gpu < len(ids) and ids[self.gpu].isdigit() else str(self.gpu)
        self.thread = None
        self.period = float(os.environ.get("XR_BENCH_CLOCK_PERIOD_MS", "2")) / 1e3
        if self._start_nvml(int(phys)):
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={phys}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.thread is not None:
            self._stop_flag = True
            self.thread.join(timeout=1.0)
        elif self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            time.sleep(0.05)
            self.proc.terminate()
        inside = [r for ts, r in self.rows if self.t0 <= ts <= self.t1]
        window = "timed region"
        if not inside:  # region shorter than the sampling period: use the whole loaded run
            inside, window = [r for _, r in self.rows], "warm-up + timed region"
        sm, mx, reasons = [], [], set()
        for r in inside:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


def bind_to_gpu_numa_node(local_rank: int) -> None:
    """Run this rank (and allocate its pinned host buffers) on the CPUs NVML reports as local to
    its GPU: with 8 ranks per box the host->device copies otherwise cross the socket link."""
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [x.strip() for x in vis.split(",")] if vis else []
        phys = int(ids[local_rank]) if local_rank < len(ids) and ids[local_rank].isdigit() else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:   # no NVML / not permitted: keep the inherited affinity
        pass


def reference_arm(args, rank, world):
    """The reference's CPU path for the same workload, all host threads (rank 0 only)."""
    if rank != 0:
        return
    import torch

    from oracle import cpu_baseline
    from xfmr_rec_b200.data import synthetic_batch

    batch = synthetic_batch(N_ITEMS, BATCH, SEQ_LEN, dim=DIM, seed=0)
    cores = os.cpu_count() or 1
    # every step is the FULL batch; the run is bounded in wall time (a few minutes whatever K is):
    # the number of steps actually timed is reported in `sample`
    budget_s, t_start = 150.0, time.perf_counter()
    times = cpu_baseline.time_train_steps(batch, 1, warmup=max(1, min(args.warmup, 2)))
    while len(times) < args.steps and time.perf_counter() - t_start + times[-1] < budget_s:
        times += cpu_baseline.time_train_steps(batch, 1, warmup=0)
    ms = 1e3 * sum(times) / len(times)
    value = BATCH / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "seq/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_dict(world),
        "cpu_baseline": {"value": value, "unit": "seq/s", "cores": cores, "kind": "port",
                         "sample": f"{len(times)} full steps (B={BATCH}) of the same workload timed "
                                   f"(of {args.steps} requested; 150 s wall budget), reference-lean form, "
                                   "torch CPU ops on the host cores, fp32"},
        "host_device": "cpu (no GPU is used by this arm; rank 0 only, one batch whatever N is)",
        "e2e": {"value": value, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "torch_threads": torch.get_num_threads(),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-retrieval", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[2] / configs[4] / HBM-kernel legs")
    ap.add_argument("--catalog", type=int, default=10_000_000)
    ap.add_argument("--cpu-steps", type=int, default=3)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import xfmr_rec_b200 as xr
    from xfmr_rec_b200.data import synthetic_batch

    all_cpus = os.sched_getaffinity(0)
    if world > 1 and not os.environ.get("XR_BENCH_NO_BIND"):
        bind_to_gpu_numa_node(local_rank)   # pinned staging buffers land next to this rank's GPU
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: ONE JSON line
        # the per-step collectives run UNDER the next step's kernels: the persistent tensor-core kernel
        # leaves `reserve` SMs free for NCCL's CTAs (dist.reserve_sms) and NCCL is told to use that many
        # Measured on 2 / 4 / 8 B200 (profiles/README.md): 8 SMs suffice where NCCL reduces in the switch (8 GPUs:
        # NVLS) or with one peer (2 GPUs); the 4-GPU ring needs 16 CTAs to move 16 MB inside one step (0.388 ms
        # per step with 8, 0.295 with 16, 0.336 with no reservation); at 8 GPUs 4 are too few (0.333 against 0.296)
        reserve = int(os.environ.get("XR_BENCH_RESERVE_SMS", "16" if world == 4 else "8"))
        if reserve > 0:
            os.environ.setdefault("NCCL_MAX_CTAS", str(reserve))
        dist.init_process_group("nccl", device_id=dev)
        xr.dist.reserve_sms(reserve)
    peaks = load_peaks()
    lib = xr._native.lib()

    # ---- synthetic inputs, weak scaling ---------------------------------------------
    # every rank gets the SAME sequence-length profile (so M, M_a and therefore the work per rank
    # are identical: weak scaling measures the machine, not the luck of the length draw) but its
    # own items and encoder outputs: a rank-specific permutation of the item ids and fresh tokens
    batch = synthetic_batch(N_ITEMS, BATCH, SEQ_LEN, dim=DIM, seed=0)
    if rank > 0:
        rng = np.random.default_rng(1000 + rank)
        perm = np.concatenate([[0], rng.permutation(N_ITEMS) + 1]).astype(np.int64)
        for key in ("history_item_idx", "pos_item_idx", "neg_item_idx"):
            batch[key] = perm[batch[key]]
        batch["token_embeddings"] = (rng.standard_normal(batch["token_embeddings"].shape)
                                     / math.sqrt(DIM)).astype(np.float32)
    emb = xr.models.ItemEmbeddings(torch.from_numpy(batch["table"]), add_padding_row=False).to(dev)
    emb.weight_bf16(), emb.rownz()
    host = {k: torch.from_numpy(batch[k]).pin_memory() for k in
            ("history_item_idx", "pos_item_idx", "neg_item_idx")}
    host_tok = torch.from_numpy(batch["token_embeddings"]).bfloat16().pin_memory()
    d_idx = {k: v.to(dev) for k, v in host.items()}
    d_tok = host_tok.to(dev)
    loss_fn = xr.InfoNCELoss(xr.LossConfig())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step_modular():
        """The drop-in modules (compute_embeds + loss + backward), one host sync for the counts."""
        tok = d_tok.detach().requires_grad_(True)
        out = xr.models.compute_embeds(emb, tok, d_idx["history_item_idx"], d_idx["pos_item_idx"],
                                       d_idx["neg_item_idx"], candidate_dtype=torch.bfloat16)
        loss = loss_fn(out["query_embed"], out["candidate_embed"])
        loss.backward()
        return loss, tok.grad, out

    # the same kernels as ONE sync-free, graph-replayed call (xr_pool_step); two objects alternate
    # so that consecutive steps never reuse a buffer and the next H2D copy overlaps the kernels
    steps2 = [xr.PoolLossStep(emb, loss_fn, BATCH, SEQ_LEN, token_dtype=torch.bfloat16) for _ in range(2)]
    host_args = (host_tok, host["history_item_idx"], host["pos_item_idx"], host["neg_item_idx"])
    for st in steps2:   # device-resident inputs for the `value` leg
        st.load(d_tok, d_idx["history_item_idx"], d_idx["pos_item_idx"], d_idx["neg_item_idx"])
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    per_rank = {}

    def max_over_ranks(ms, tag=None):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            if tag:
                allt = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(allt, t)
                per_rank[tag] = [round(float(x), 4) for x in allt]
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        return ms

    # ---- data-parallel collectives of a real training step (N > 1): the all-reduce(sum) of the loss for
    #      logging (dist.reduce_loss: the reference's losses are sums over rows) and the all-reduce of the
    #      encoder gradients (<= 16 MB fp32 for the 2-layer d=384 encoder), issued every step on
    #      NCCL's own stream so that they overlap the next step's kernels -------------------------------
    grad_buf = torch.zeros(4 << 20, dtype=torch.float32, device=dev) if world > 1 else None   # 16 MB
    pending = []

    def dp_collectives(loss):
        if world == 1:
            return
        while len(pending) > 2:          # keep at most two steps of collectives in flight
            pending.pop(0).wait()
        l = loss.detach().clone()
        pending.append(dist.all_reduce(l, op=dist.ReduceOp.SUM, async_op=True))
        pending.append(dist.all_reduce(grad_buf, op=dist.ReduceOp.SUM, async_op=True))

    def timed_resident(steps, warmup, with_collectives):
        """value: inputs resident in HBM.  K graph replays queued back to back (the two buffer sets
        alternate, so a step never finds its inputs in L2), ONE CUDA-event pair around the region:
        the host runs ahead of the device, as it does inside a training loop."""
        sampler = ClockSampler(local_rank)
        sampler.start()
        for i in range(warmup):
            l, _ = steps2[i % 2].run()
            if with_collectives:
                dp_collectives(l)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        last = None
        sampler.mark_begin()
        a.record()
        for i in range(steps):
            last = steps2[i % 2].run()
            if with_collectives:
                dp_collectives(last[0])
        while pending:
            pending.pop(0).wait()
        b.record()
        barrier()
        sampler.mark_end()
        clocks = sampler.stop()
        return max_over_ranks(a.elapsed_time(b), "value_region_ms" if with_collectives or world == 1 else None), clocks, last

    def timed_e2e(steps, warmup, tokens_on_device=False):
        """e2e: every step copies its inputs from pinned host memory and reads the loss back.
        The H2D copy of step i+1 is enqueued (copy stream) before the host blocks on the loss of
        step i, so it overlaps step i's kernels; timed as one region over all K steps."""
        host_loss = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        read_ev = [torch.cuda.Event() for _ in range(2)]
        src = (d_tok,) + host_args[1:] if tokens_on_device else host_args

        def loop(n):
            vals = []
            steps2[0].load(*src)
            for i in range(n):
                loss, _ = steps2[i % 2].run()                        # compute phase of batch i
                host_loss[i % 2].copy_(loss.reshape(1), non_blocking=True)   # device -> host read of its result
                read_ev[i % 2].record()
                dp_collectives(loss)
                if i + 1 < n:
                    steps2[(i + 1) % 2].load(*src)                   # H2D copies + ingest of batch i+1
                if i >= 1:                                           # the host consumes loss i-1 while step i runs
                    read_ev[(i - 1) % 2].synchronize()
                    vals.append(float(host_loss[(i - 1) % 2]))
            read_ev[(n - 1) % 2].synchronize()
            vals.append(float(host_loss[(n - 1) % 2]))
            while pending:
                pending.pop(0).wait()
            return vals
        loop(warmup)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        vals = loop(steps)
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b), None if tokens_on_device else "e2e_region_ms"), vals

    def timed_modular(steps, warmup, profile):
        for _ in range(warmup):
            step_modular()
        barrier()
        if profile:
            lib.xr_fused_profile(1)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(steps)]
        last = None
        for a, b in evs:
            flush.fill_(1)
            a.record()
            last = step_modular()
            b.record()
        barrier()
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in evs)), last

    # ---- device-resident throughput (with the DP collectives in the region when N > 1) ------------
    total_ms, clocks, last = timed_resident(args.steps, args.warmup, with_collectives=True)
    loss_val = float(last[0])
    m_a, m_rows = steps2[0].row_counts()
    c_cols = m_a + 1
    ms_per_step = total_ms / args.steps
    value = world * BATCH / (ms_per_step / 1e3)
    no_coll_ms = None
    if world > 1:    # the same region without any collective: what round 1 reported as `value`
        no_coll_ms, _, _ = timed_resident(args.steps, 3, with_collectives=False)
        no_coll_ms /= args.steps

    # ---- end to end: host buffers in, loss scalar out --------------------------------------------
    e2e_ms, e2e_vals = timed_e2e(args.steps, args.warmup)
    e2e_value = world * BATCH / (e2e_ms / args.steps / 1e3)
    h2d = sum(v.numel() * v.element_size() for v in host.values()) + host_tok.numel() * 2
    d2h = 4
    assert abs(e2e_vals[-1] - loss_val) <= 1e-6 * abs(loss_val), (e2e_vals[-1], loss_val)
    # the same loop when the encoder output is already on the device, as it is in real training (the
    # encoder runs on the GPU): only the three index tensors cross the host link
    e2e_dev_ms, _ = timed_e2e(args.steps, 3, tokens_on_device=True)
    h2d_idx = sum(v.numel() * v.element_size() for v in host.values())

    # ---- the trainer's whole compute_losses (trainer.py:213-264): train loss forward/backward AND
    #      LogitsStatistics + all seven losses, one sync-free graph replay per step ----------------
    mon_steps = min(args.steps, 50)
    mon = xr.PoolLossStep(emb, loss_fn, BATCH, SEQ_LEN, token_dtype=torch.bfloat16, monitor=True)
    mon.load(d_tok, d_idx["history_item_idx"], d_idx["pos_item_idx"], d_idx["neg_item_idx"])
    for _ in range(3):
        mon.run()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(mon_steps)]
    for a, b in ev:
        flush.fill_(1)
        a.record()
        mon.run()
        b.record()
    barrier()
    mon_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev)) / mon_steps
    mon_losses, mon_stats = mon.loss_dict()
    assert abs(float(mon_losses["loss/InfoNCELoss"]) - loss_val) <= 1e-4 * abs(loss_val), "monitor pass and train loss disagree"
    del mon

    # ---- the fused kernel alone (roofline leg) + the drop-in module path ------------------------
    # CUDA events recorded inside the library around every launch of the main fused kernel, on the
    # launching stream, over a timed loop of the SAME step issued without graph capture, L2 flushed
    mod_steps = min(args.steps, 50)
    mod_ms, mod_last = timed_modular(mod_steps, 3, profile=True)
    buf = (ctypes.c_float * 512)()
    n_prof = lib.xr_fused_profile_read(buf, 512)
    lib.xr_fused_profile(0)
    kern_ms = [buf[i] for i in range(max(n_prof, 0))]
    assert float(mod_last[0]) == loss_val, "graph-replayed step and module path disagree"

    flops = 4.0 * m_rows * c_cols * DIM          # scores + dQ, table frozen
    k_ms = statistics.median(kern_ms) if kern_ms else float("nan")
    achieved = flops / (k_ms * 1e-3) / 1e12 if kern_ms else float("nan")
    line = {
        "metric": METRIC, "value": value, "unit": "seq/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": config_dict(world),
        "detail": {"rows_M": m_rows, "candidates_C": c_cols,
                   "l2": "inputs larger than L2: two alternating buffer sets, ~2 x 125 MB touched per "
                         "pair of steps > 126 MB L2 (value and e2e); the kernel-only roofline leg "
                         "flushes L2 with a 256 MB write between launches",
                   "api": "PoolLossStep (xr_pool_step, CUDA-graph replay)",
                   "parallelism": f"dp{world} (independent batches, table replicated)" + (
                       "; per step INSIDE the timed region: all-reduce(sum) of the loss (dist.reduce_loss) "
                       "and a 16 MB fp32 all-reduce standing in for the encoder gradients, on NCCL's stream, "
                       "overlapping the next step" if world > 1 else "")},
        "e2e": {"value": e2e_value, "unit": "seq/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                "note": "every step starts from pinned HOST inputs (H2D copy of batch i+1 on the copy stream "
                        "under the kernels of batch i); the loss of every step is copied to pinned host "
                        "memory and read by the host one step later (while the next step runs), as a "
                        "training loop logs it.  Bound by the host link (20.3 MB per step)",
                "encoder_output_on_device": {
                    "value": world * BATCH / (e2e_dev_ms / args.steps / 1e3), "unit": "seq/s",
                    "ms_per_step": e2e_dev_ms / args.steps, "h2d_bytes_per_step": h2d_idx,
                    "note": "same loop with the encoder-output stand-in resident on the device (where a real "
                            "training step leaves it): only the three index tensors cross the host link"}},
        "gpu_launches": 7 * args.steps,   # xr_pool_step: compaction, plan, gather, diagonal, fused, finalize, row sum
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "fused_pool_kernel<InfoNCE> (tcgen05, stream-K)",
                     "achieved": achieved, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["tf_burst"] if kern_ms else None,
                     "frac_of_sustained_peak": achieved / peaks["tf_sustained"] if kern_ms else None,
                     "peak_sustained": peaks["tf_sustained"],
                     # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture
                     # of this kernel on this workload (profiles/ncu_fused_pool_kernel_r02_raw.txt: 19.10 MB
                     # read + 1.41 MB written; algorithmic operand bytes 28.3 MB; partial dQ stays in L2)
                     "traffic": 20.5e6, "traffic_unit": "bytes",
                     "peak_source": f"{peaks['src']}: burst bf16 (the kernel is timed alone, L2 flushed); "
                                    "sustained figure beside it",
                     "kernel_ms": k_ms, "kernel_ms_min": min(kern_ms) if kern_ms else None,
                     "kernel_share_of_step": k_ms / ms_per_step if kern_ms else None,
                     "algorithmic_flops_per_launch": flops},
        "loss": loss_val,
        "per_rank_region_ms": per_rank or None,
        "compute_losses": {"value": world * BATCH / (mon_ms / 1e3), "unit": "seq/s", "ms_per_step": mon_ms,
                           "gpu_launches_per_step": 11,
                           "note": "PoolLossStep(monitor=True): the step above PLUS LogitsStatistics and all "
                                   "seven losses (what trainer.py:250-263 logs every step) from the SAME "
                                   "tensor-core pass (xr_pool_step_compute_mon: the train kernel accumulates "
                                   "the monitoring sums of both logit families), one graph replay; L2 flushed "
                                   "between steps",
                           "losses": {k: float(v) for k, v in mon_losses.items()}},
        "module_api": {"value": world * BATCH / (mod_ms / mod_steps / 1e3), "unit": "seq/s",
                       "ms_per_step": mod_ms / mod_steps,
                       "note": "same step through compute_embeds + InfoNCELoss + backward (the "
                               "reference's call sequence; one host sync for the row counts)"},
    }
    if world > 1:
        xr.dist.reserve_sms(0)      # the retrieval leg has no collective running beside its kernels
        line["dp_collectives"] = {
            "reserved_sms": reserve,
            "per_step": ["all_reduce(sum) loss scalar (dist.reduce_loss)", "all_reduce(sum) 16 MB fp32 (encoder-gradient sized)"],
            "ms_per_step_with": ms_per_step, "ms_per_step_without": no_coll_ms,
            "value_without": world * BATCH / (no_coll_ms / 1e3)}
    del steps2
    torch.cuda.empty_cache()

    if not args.no_extra:
        line["configs"] = {}
        line["configs"]["cfg3_ccl_sampled"] = bench_cfg3(xr, dev, world, peaks, max_over_ranks, flush)
        line["configs"]["cfg5_points"] = bench_cfg5(xr, dev, peaks, flush) if rank == 0 or world == 1 else None
        line["configs"]["hbm_kernels"] = bench_hbm_kernels(xr, dev, peaks, flush) if rank == 0 or world == 1 else None
        line["configs"]["encoder_train_step"] = bench_encoder_step(xr, dev, flush) if rank == 0 or world == 1 else None
        torch.cuda.empty_cache()
    if not args.no_retrieval:
        line["retrieval"] = bench_retrieval(args, xr, dev, rank, world, peaks)
    # every GPU leg is done: leave the process group BEFORE the CPU legs (only rank 0 runs them; they use
    # all host cores, and an earlier version that ran them before the retrieval leg made rank 0 stall
    # for tens of milliseconds inside the timed searches while the other ranks waited in the exchange)
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        # ---- CPU baseline: same workload, reference arithmetic on the host cores ------------------
        from oracle import cpu_baseline

        for tid in os.listdir("/proc/self/task"):   # the CPU leg uses every host core again
            try:                                       # (worker threads inherited the NUMA mask)
                os.sched_setaffinity(int(tid), all_cpus)
            except OSError:
                pass
        b0 = synthetic_batch(N_ITEMS, BATCH, SEQ_LEN, dim=DIM, seed=0)
        times = cpu_baseline.time_train_steps(b0, args.cpu_steps, warmup=1)
        cpu_ms = 1e3 * sum(times) / len(times)
        if world == 1:   # the trainer's whole compute_losses on the host cores (one full step)
            cl_times, _ = cpu_baseline.time_compute_losses(b0, 1, warmup=0)
            line["compute_losses"]["cpu_baseline"] = {
                "value": BATCH / cl_times[0], "unit": "seq/s", "ms_per_step": 1e3 * cl_times[0],
                "cores": os.cpu_count(), "kind": "port",
                "sample": "1 full step (B=%d): LogitsStatistics + all seven losses + InfoNCE backward, "
                          "reference-lean torch CPU ops, fp32" % BATCH}
        line["cpu_baseline"] = {"value": BATCH / (cpu_ms / 1e3), "unit": "seq/s",
                                "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{args.cpu_steps} full steps (B={BATCH}) of the same "
                                          "workload, reference-lean torch CPU ops, fp32",
                                "ms_per_step": cpu_ms}
        if not args.no_retrieval and world == 1:   # exact cosine top-k on the host cores, bounded sample
            rows_s, q_s = min(args.catalog, 1_000_000), 256
            dt = cpu_baseline.time_exact_search(rows_s, q_s, 100, dim=DIM)
            line["retrieval"]["cpu_baseline"] = {
                "value": q_s / dt * rows_s / args.catalog, "unit": "queries/s", "cores": os.cpu_count(),
                "kind": "port",
                "sample": f"{q_s} queries x {rows_s} fp32 rows (chunked matmul + stable sort, torch CPU ops) in "
                          f"{dt:.2f} s = {q_s / dt:.1f} queries/s at that size, scaled linearly in the "
                          f"number of rows to the {args.catalog}-row catalog"}
        print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def _timeit(torch, fn, flush, reps=10, warm=3):
    """Median of `reps` CUDA-event timings, L2 flushed before each."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        if flush is not None:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return statistics.median(out)


def bench_cfg3(xr, dev, world, peaks, max_over_ranks, flush):
    """BASELINE configs[2]: CCL (AlignmentContrastiveLoss, cosine) with 512 sampled negatives per positive,
    87,585-item table, global batch 1,024 sequences x L = 50 data-parallel over 8 GPUs = 6,400 rows per GPU
    (weak scaling: every rank runs 6,400 rows).  One step = the one-pass gather-dot kernel
    (xr_sampled_step: logits + EmbedLoss pipeline + dL/dq) + backward glue through the loss module."""
    import torch

    from xfmr_rec_b200 import ops

    n_items, d, k, m = 87_585, DIM, 512, 6400
    g = torch.Generator(device=dev).manual_seed(7)
    table = (torch.randn((n_items + 1, d), generator=g, device=dev) / d ** 0.5).bfloat16()
    table[0] = 0
    _, table_inv = ops.normalize_rows(table, 1e-8, want_y=False)
    q = (torch.randn((m, d), generator=g, device=dev) / d ** 0.5).bfloat16()
    idx = torch.randint(1, n_items + 1, (m, k + 1), generator=g, device=dev)
    cand = xr.SampledCandidates(table, idx, table_inv)
    fn = xr.AlignmentContrastiveLoss(xr.LossConfig())

    def step():
        qq = q.detach().requires_grad_(True)
        loss = fn(qq, cand)
        loss.backward()
        return loss

    ms = max_over_ranks(_timeit(torch, step, flush, reps=20))
    cfg = ops.make_cfg(xr.LossConfig())
    q_inv = ops.normalize_rows(q, 1e-8, want_y=False)[1]
    kind = xr._native.LOSS_KIND["AlignmentContrastiveLoss"]
    k_ms = _timeit(torch, lambda: ops.sampled_step(q, table, idx, cfg, table_inv, q_inv, kind), flush, reps=20)
    byts = m * (k + 1) * d * 2.0          # every candidate row read once for the logits
    return {"workload": "ML-32M-shaped synthetic (87,585 items), CCL, 512 sampled negatives per positive, "
                        "6,400 rows per GPU (= 1,024 sequences x L=50 over 8 GPUs), bf16 (BASELINE configs[2])",
            "metric": "rows/s (positions)", "value": world * m / (ms / 1e3), "unit": "rows/s",
            "seq_per_s": world * m / 50 / (ms / 1e3), "ms_per_step": ms, "rows_per_gpu": m, "candidates_per_row": k + 1,
            "scaling": "weak (no data-path collective)",
            "roofline": {"bound": "hbm", "kernel": "sampled_step384_kernel (gather-dot + pipeline + dL/dq, one pass)",
                         "achieved": byts / (k_ms * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                         "frac": byts / (k_ms * 1e-3) / 1e9 / peaks["hbm"], "kernel_ms": k_ms, "traffic": None,
                         "algorithmic_bytes_per_launch": byts,
                         "note": "the 67 MB bf16 table is L2-resident, so the gather-dot runs above the HBM copy "
                                 "bandwidth; reported against HBM because that is the roofline SURVEY 8d names"}}


def bench_cfg5(xr, dev, peaks, flush):
    """BASELINE configs[4]: two points of the BPR / SSM sweep, fused epilogue against the reference's logits
    materialisation (Q @ C^T -> (B, N) tensor -> the reference's loss()), both on this GPU."""
    import torch
    import torch.nn.functional as F

    from xfmr_rec_b200 import ops

    out = []
    g = torch.Generator(device=dev).manual_seed(11)
    for name, kw, (m, cn) in (("InfoNCELoss", {}, (8192, 100_000)), ("PairwiseLogisticLoss", {"margin": 0.0}, (2048, 1_000_000))):
        d = DIM
        q = (torch.randn((m, d), generator=g, device=dev) / d ** 0.5).bfloat16()
        pos = (torch.randn((m, d), generator=g, device=dev) / d ** 0.5).bfloat16()
        neg = (torch.randn((cn, d), generator=g, device=dev) / d ** 0.5).bfloat16()
        cfg = ops.make_cfg(xr.LossConfig(**kw), logits_bf16=True)
        kind = xr._native.LOSS_KIND[name]
        lib = xr._native.lib()
        for _ in range(2):
            ops.fused_pool_loss(q, pos, neg, kind, cfg)
        torch.cuda.synchronize()
        lib.xr_fused_profile(1)
        ms = _timeit(torch, lambda: ops.fused_pool_loss(q, pos, neg, kind, cfg), flush, reps=5, warm=0)
        buf = (ctypes.c_float * 64)()
        cnt = lib.xr_fused_profile_read(buf, 64)
        lib.xr_fused_profile(0)
        k_ms = statistics.median([buf[i] for i in range(cnt)]) if cnt > 0 else float("nan")

        def materialised():
            qq = q.detach().requires_grad_(True)
            logits = torch.cat([(qq * pos).sum(-1, keepdim=True), qq @ neg.T], 1)   # bf16, losses.py:195 under autocast
            keep = logits < logits[:, :1]
            if name == "InfoNCELoss":
                z = logits.float().masked_fill(~keep, float("-inf"))
                z[:, 0] = logits[:, 0].float()
                loss = F.cross_entropy(z, torch.zeros(m, dtype=torch.long, device=dev), reduction="sum")
            else:
                w = keep.float()
                loss = ((F.softplus(logits.float() - logits[:, :1].float()) * w).sum(-1) / (w.sum(-1) + 1e-9)).sum()
            loss.backward()
            return loss
        ms_mat = _timeit(torch, materialised, flush, reps=3, warm=1)
        flops = 4.0 * m * (cn + 1) * d
        out.append({"loss": name, "queries": m, "candidates": cn, "fused_ms": ms, "fused_kernel_ms": k_ms,
                    "materialised_logits_ms": ms_mat, "speedup": ms_mat / ms,
                    "roofline": {"bound": "tensor", "achieved": flops / (k_ms * 1e-3) / 1e12, "peak": peaks["tf_burst"],
                                 "unit": "TFLOP/s", "frac": flops / (k_ms * 1e-3) / 1e12 / peaks["tf_burst"],
                                 "frac_of_sustained_peak": flops / (k_ms * 1e-3) / 1e12 / peaks["tf_sustained"],
                                 "traffic": None}})
        del q, pos, neg
        torch.cuda.empty_cache()
    return {"workload": "BPR / SSM sweep points, fused epilogue vs materialised (B, N) logits on the same GPU "
                        "(BASELINE configs[4])", "points": out}


def bench_encoder_step(xr, dev, flush):
    """the train step WITH the sequence encoder (models.py:51-102, 306-345) at the configs[1]
    shape -- encoder forward (2 layers, d = 384, 12 heads, intermediate 1536 = all-MiniLM-L6-v2's, bf16-mixed)
    -> the scoring-and-loss step of the headline line -> encoder backward, all in one CUDA graph."""
    import torch

    from xfmr_rec_b200.data import synthetic_batch
    from xfmr_rec_b200.encoder import EncoderConfig, GraphedEncoderStep, SeqEncoder

    b = synthetic_batch(N_ITEMS, BATCH, SEQ_LEN, dim=DIM, seed=5)
    table = torch.from_numpy(b["table"]).to(dev)
    hist, pos, neg = (torch.from_numpy(b[k]).to(dev) for k in ("history_item_idx", "pos_item_idx", "neg_item_idx"))
    torch.manual_seed(0)
    cfg = EncoderConfig(num_hidden_layers=2, intermediate_size=1536, max_seq_length=SEQ_LEN)
    enc = SeqEncoder(cfg, compute_dtype=torch.bfloat16).to(dev).train()    # dropout 0.1 / 0.1 (HF defaults)
    emb = xr.models.ItemEmbeddings(table, add_padding_row=False).to(dev)
    step = xr.PoolLossStep(emb, xr.InfoNCELoss(xr.LossConfig()), BATCH, SEQ_LEN, token_dtype=torch.float32,
                           logits_bf16=True, use_graph=False)
    graphed = GraphedEncoderStep(enc, step, table, SEQ_LEN)
    ms = _timeit(torch, lambda: graphed(hist, pos, neg), flush, reps=20)
    # the same with the optimizer update inside the graph (what a Lightning training_step + optimizer_step does)
    del graphed
    enc.zero_grad(set_to_none=True)
    trained = [p for n, p in enc.named_parameters() if not n.startswith(("pooler", "embeddings.word"))]
    opt = torch.optim.AdamW(trained, lr=1e-3, weight_decay=0.01, capturable=True, fused=True)   # config.yaml: lr, wd
    step2 = xr.PoolLossStep(emb, xr.InfoNCELoss(xr.LossConfig()), BATCH, SEQ_LEN, token_dtype=torch.float32,
                            logits_bf16=True, use_graph=False)
    graphed = GraphedEncoderStep(enc, step2, table, SEQ_LEN, optimizer=opt)
    ms_opt = _timeit(torch, lambda: graphed(hist, pos, neg), flush, reps=20)
    with torch.no_grad():
        fwd_ms = _timeit(torch, lambda: enc.encode_tokens(hist, table), flush, reps=10)
    loss = float(graphed(hist, pos, neg))
    n_par = sum(p.numel() for n, p in enc.named_parameters() if p.grad is not None)
    return {"value": BATCH / (ms / 1e3), "unit": "seq/s", "ms_per_step": ms, "ms_per_step_with_adamw": ms_opt,
            "encoder_forward_eager_ms": fwd_ms,
            "loss": loss, "trained_parameters": n_par,
            "encoder": {"layers": 2, "hidden": DIM, "heads": 12, "intermediate": 1536, "compute": "bf16-mixed "
                        "(bf16 GEMMs / attention / GELU, fp32 residual stream, LayerNorm and softmax)",
                        "dropout": "0.1 hidden / 0.1 attention (training mode, HF defaults): counter-based masks "
                                   "recomputed in the backward"},
            "note": "GraphedEncoderStep: encoder forward + PoolLossStep + encoder backward (parameter gradients "
                    "in .grad) replayed as one CUDA graph, and the same graph with the AdamW update inside "
                    "(ms_per_step_with_adamw); L2 flushed between steps; the linear "
                    "layers are cuBLAS GEMMs, everything else this repository's kernels (csrc/encoder.cu). "
                    "Reference encoder class (transformers BertModel, eager, autocast bf16) on the same GPU: "
                    "profiles/encoder_r02.json"}


def bench_hbm_kernels(xr, dev, peaks, flush):
    """HBM rooflines of the gather and top-k families, measured in this run (north_star: >= 70 %)."""
    import torch

    from xfmr_rec_b200 import ops

    g = torch.Generator(device=dev).manual_seed(3)
    rows, n_table = 1 << 20, 4_000_000
    table = torch.randn((n_table, DIM), generator=g, device=dev).bfloat16()      # 3 GB: far larger than L2
    idx = torch.randint(0, n_table, (rows,), generator=g, device=dev)
    ms = _timeit(torch, lambda: ops.gather_rows(table, idx), flush, reps=10)
    gb = 2.0 * rows * DIM * 2 + rows * 8
    out = {"gather_rows": {"rows": rows, "table_rows": n_table, "dtype": "bf16", "kernel_ms": ms,
                           "roofline": {"bound": "hbm", "achieved": gb / (ms * 1e-3) / 1e9, "peak": peaks["hbm"],
                                        "unit": "GB/s", "frac": gb / (ms * 1e-3) / 1e9 / peaks["hbm"], "traffic": None,
                                        "algorithmic_bytes_per_launch": gb}}}
    del table
    torch.cuda.empty_cache()
    u, n, k = 64, 10_000_000, 100
    sc = torch.randn((u, n), generator=g, device=dev)
    ms = _timeit(torch, lambda: ops.topk(sc, k), None, reps=6)     # 2.56 GB of scores: larger than L2
    gb = u * n * 4.0 + u * k * 12
    out["topk"] = {"rows": u, "columns": n, "k": k, "ms": ms,
                   "roofline": {"bound": "hbm", "achieved": gb / (ms * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                                "frac": gb / (ms * 1e-3) / 1e9 / peaks["hbm"], "traffic": None,
                                "algorithmic_bytes_per_launch": gb,
                                "note": "xr_topk over a materialised (64, 10M) fp32 score matrix (stream + merge + emit)"}}
    return out


def bench_retrieval(args, xr, dev, rank, world, peaks):
    """BASELINE configs[3]: full-catalog top-100 queries/s, 10M x 384 bf16 catalog sharded by rows across
    the ranks, 20-200 excluded ids per query (the user's history, index.py:239-247), U in {1, 256, 4096};
    per-shard top-k merged through NVLink peer memory (or an NCCL all-gather).  The local search is one
    CUDA-graph replay of xr_score_topk (sample thresholds -> filter in the tcgen05 scoring epilogue ->
    exact top-k of the survivors)."""
    import gc

    import torch
    import torch.distributed as dist

    from xfmr_rec_b200 import ops
    from xfmr_rec_b200.dist import ShardedIndex, shard_range

    lib = xr._native.lib()
    n, k = args.catalog, 100
    lo, hi = shard_range(n, rank, world)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    shard = torch.empty((hi - lo, DIM), dtype=torch.bfloat16, device=dev)
    for a in range(0, hi - lo, 1_000_000):
        shard[a:a + 1_000_000] = torch.randn((min(1_000_000, hi - lo - a), DIM), generator=g, device=dev).bfloat16()
    idx = xr.index.ExactIndex(xr.index.ExactIndexConfig(index_metric="cosine", dtype="bf16"), dev, row_offset=lo)
    idx.catalog, _ = ops.normalize_rows(shard, 1e-12, torch.bfloat16)
    del shard
    sharded = ShardedIndex(idx, exchange="auto" if world > 1 else "nccl", use_plan=True, plan_max_exclusions=200)
    points = []
    for u in (256, 1, 4096):
        gq = torch.Generator(device=dev).manual_seed(99 + u)
        q = torch.randn((u, DIM), generator=gq, device=dev)
        lens = torch.randint(20, 201, (u,), generator=gq, device=dev)
        offs = torch.zeros(u + 1, dtype=torch.int64, device=dev)
        offs[1:] = lens.cumsum(0)
        rows = torch.randint(0, n, (int(offs[-1]),), generator=gq, device=dev)
        excl = (offs, rows)
        reps = 12 if u <= 256 else 4
        for _ in range(3):
            s, i = sharded.search_batch(q, excl, k, check=False)
        torch.cuda.synchronize()
        gc.collect()
        gc.disable()    # a cyclic-GC pass of the host interpreter inside a 0.5 ms search shows up as a 5 ms search
        if world > 1:
            dist.barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in evs:
            a.record()
            s, i = sharded.search_batch(q, excl, k, check=False)
            b.record()
        torch.cuda.synchronize()
        gc.enable()
        assert not sharded.plans_overflowed(), "a survivor list overflowed: the timed searches are not exact"
        in_order = [a.elapsed_time(b) for a, b in evs]
        ms = statistics.median(in_order)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        # the scoring kernel alone: a few eagerly issued local searches with the in-library event hook
        lib.xr_fused_profile(1)
        for _ in range(3):
            idx.search_batch(q, excl, k, max_exclusions=200)
        torch.cuda.synchronize()
        buf = (ctypes.c_float * 64)()
        cnt = lib.xr_fused_profile_read(buf, 64)
        lib.xr_fused_profile(0)
        evt = sorted(buf[j] for j in range(cnt))
        k_ms = evt[-2] if cnt >= 2 else float("nan")     # per search: [sample pass, filter pass]; the filter pass is the long one
        flops, byts = 2.0 * u * (hi - lo) * DIM, (hi - lo) * DIM * 2.0
        tensor_bound = u >= 211
        ach = flops / (k_ms * 1e-3) / 1e12 if tensor_bound else byts / (k_ms * 1e-3) / 1e9
        peak = peaks["tf_sustained"] if tensor_bound else peaks["hbm"]
        points.append({
            "queries": u, "exclusions_per_query": "20-200", "ms_per_batch": ms, "queries_per_s": u / (ms / 1e3),
            "ms_per_batch_mean": sum(in_order) / reps, "ms_per_batch_min_max": [min(in_order), max(in_order)],
            "ms_per_search_in_order": [round(x, 4) for x in in_order],
            "limiting_kernel": "score_gmax2_kernel<FILTER> (tcgen05 cta_group::2)" if u > 128 else "fused_pool_kernel<FILTER> (tcgen05)",
            "roofline": {"bound": "tensor" if tensor_bound else "hbm", "achieved": ach, "peak": peak,
                         "unit": "TFLOP/s" if tensor_bound else "GB/s", "frac": ach / peak, "kernel_ms": k_ms,
                         "kernel_share_of_search": k_ms / ms, "traffic": None,
                         "note": "2*U*N*D FLOP against the sustained bf16 peak (the kernel runs for milliseconds "
                                 "under the power cap)" if tensor_bound else
                                 "catalog bytes N*D*2 read once against the measured HBM COPY bandwidth (a copy reads and "
                                 "writes; a read-only stream can exceed it, so frac may pass 1)"}})
    head = points[0]
    out = {"metric": "full-catalog top-100 queries/sec", "value": head["queries_per_s"], "unit": "queries/s",
           "catalog_rows": n, "queries": head["queries"], "k": k, "ms_per_batch": head["ms_per_batch"],
           "exclusions_per_query": "20-200 (CSR through the plan's static buffers)",
           "timing": "median of 12 searches (4 at U=4096), CUDA events per search, max over ranks",
           "dtype": "bf16", "catalog_bytes_per_rank": (hi - lo) * DIM * 2,
           "path": "xr_score_topk as one CUDA-graph replay: strided-sample group maxima -> (k+28)-th largest = "
                   "threshold -> tcgen05 scoring pass with the filter in the epilogue -> exact top-k of the "
                   "survivors (+ exchange of (U,k) when sharded)",
           "exchange": ("none (one GPU)" if world == 1 else
                        "NVLink peer memory: xr_topk_merge_peers after one device-side barrier"
                        if sharded._peer is not None else f"NCCL all-gather + merge ({sharded._peer_failed})"),
           "points": points}
    # ---- multi-GPU parity carried by the bench line: sharded search == unsharded search -------------
    if world > 1:
        n_chk, u_chk = 1 << 20, 256
        gc_ = torch.Generator(device=dev).manual_seed(4321)            # the SAME catalog on every rank
        full = torch.randn((n_chk, DIM), generator=gc_, device=dev).bfloat16()
        full[::1000] = full[7]                                          # duplicated rows: ties across shards
        qc = torch.randn((u_chk, DIM), generator=gc_, device=dev)
        ex = [torch.randint(0, n_chk, (int(x),), generator=gc_, device=dev).tolist()
              for x in torch.randint(0, 60, (u_chk,), generator=gc_, device=dev).tolist()]
        one = xr.index.ExactIndex(xr.index.ExactIndexConfig(index_metric="cosine", dtype="bf16"), dev).set_catalog(full)
        s1, i1 = one.search_batch(qc, ex, k)
        clo, chi = shard_range(n_chk, rank, world)
        part = xr.index.ExactIndex(xr.index.ExactIndexConfig(index_metric="cosine", dtype="bf16"), dev, row_offset=clo)
        part.catalog = one.catalog[clo:chi]
        s2, i2 = ShardedIndex(part, exchange="auto").search_batch(qc, ex, k)
        same = torch.tensor([int(torch.equal(i1, i2) and torch.equal(s1, s2))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        out["sharded_equals_unsharded"] = bool(int(same))
        out["sharded_equals_unsharded_check"] = (f"{u_chk} queries with 0-59 exclusions over a {n_chk}-row catalog with "
                                                 f"1,049 duplicated rows: {world} shards + exchange vs one GPU, "
                                                 "indices and scores bit-identical on every rank")
        assert out["sharded_equals_unsharded"], "sharded search differs from the unsharded search"
    return out


if __name__ == "__main__":
    main()
