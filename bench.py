#!/usr/bin/env python
"""bench.py — scoring-and-loss train step (BASELINE config 2) on N B200s.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (`config.workload`): MovieLens-20M-shaped synthetic batch — 27,278 items, 384-d frozen
item table, B=128 sequences of up to L=200 positions, InfoNCE (sampled softmax) with the
in-batch shared negative pool, bf16 tensor-core arithmetic.  One step = one pass of the hot path
over one batch: index compaction + the three embedding gathers (compute_embeds,
models.py:366-419) + fused contraction / loss / gradient (losses.py:150-155, 479-488) +
the scatter of dL/dquery back to the encoder-output layout.  The sequence encoder is outside the
path (north_star); a random (B, L, 384) tensor stands in for its output.

`value`  : sequences/s with every input already resident in HBM.
`e2e`    : the same step through the public API with HOST (pinned) inputs: per step the index
           tensors and the encoder-output stand-in are copied host->device and the loss scalar is
           read back.
`roofline`: the fused tcgen05 kernel alone, timed per launch with CUDA events on its stream,
           4*M*C*D algorithmic FLOPs (scores + dQ) against the measured sustained bf16 peak.
`cpu_baseline` / `--impl reference`: the reference's arithmetic in torch CPU ops on all host
           threads (oracle/cpu_baseline.py; the reference is Python and cannot travel).
"""

from __future__ import annotations

import argparse
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

N_ITEMS, BATCH, SEQ_LEN, DIM = 27278, int(os.environ.get("XR_BENCH_BATCH", "128")), 200, 384
WORKLOAD = (f"ML-20M-shaped synthetic (27,278 items, 384-d), B={BATCH} x L=200, InfoNCE in-batch "
            "shared-pool negatives, bf16 (BASELINE configs[1])")


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe): an
    `nvidia-smi -lms` child process (no GIL contention with the launch thread) started before the
    warm-up; samples are attributed to the timed window by their wall-clock arrival time."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None
        self.t0 = self.t1 = None

    def _start_nvml(self, phys: int) -> bool:
        """Preferred: poll NVML from a thread every 2 ms (the timed region lasts tens of ms)."""
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

    from oracle import cpu_baseline, xfmr_oracle as orc

    batch = orc.synth_batch(N_ITEMS, BATCH, SEQ_LEN, dim=DIM, seed=0)
    cores = os.cpu_count() or 1
    # every step is the FULL batch; the run is bounded in wall time (a few minutes whatever K is):
    # the number of steps actually timed is reported in `sample`
    import time

    budget_s, t_start = 150.0, time.perf_counter()
    times = cpu_baseline.time_train_steps(batch, 1, warmup=max(1, min(args.warmup, 2)))
    while len(times) < args.steps and time.perf_counter() - t_start + times[-1] < budget_s:
        times += cpu_baseline.time_train_steps(batch, 1, warmup=0)
    ms = 1e3 * sum(times) / len(times)
    value = BATCH / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "train seq/sec (scoring-and-loss step)", "value": value,
        "unit": "seq/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": BATCH, "seq_len": SEQ_LEN},
        "cpu_baseline": {"value": value, "unit": "seq/s", "cores": cores, "kind": "port",
                         "sample": f"{len(times)} full steps (B={BATCH}) of the same workload timed "
                                   f"(of {args.steps} requested; 150 s wall budget), reference-lean form, "
                                   "torch CPU ops on the host cores, fp32"},
        "host_device": "cpu (no GPU is used by this arm)",
        "e2e": {"value": value, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "torch_threads": torch.get_num_threads(),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-retrieval", action="store_true")
    ap.add_argument("--catalog", type=int, default=10_000_000)
    ap.add_argument("--queries", type=int, default=256)
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
    from oracle import xfmr_oracle as orc  # synthetic inputs + cpu_baseline leg only

    all_cpus = os.sched_getaffinity(0)
    if world > 1 and not os.environ.get("XR_BENCH_NO_BIND"):
        bind_to_gpu_numa_node(local_rank)   # pinned staging buffers land next to this rank's GPU
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    peak_tf, peak_hbm, peak_src = load_peaks()

    # ---- synthetic inputs, weak scaling ---------------------------------------------
    # every rank gets the SAME sequence-length profile (so M, M_a and therefore the work per rank
    # are identical: weak scaling measures the machine, not the luck of the length draw) but its
    # own items and encoder outputs: a rank-specific permutation of the item ids and fresh tokens
    batch = orc.synth_batch(N_ITEMS, BATCH, SEQ_LEN, dim=DIM, seed=0)
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
    # the second operand set exists so consecutive timed steps never touch the same HBM lines:
    # per-step traffic (~70 MB incl. partials) is below the 126 MB L2, so L2 is flushed between
    # timed iterations by writing a 256 MB buffer
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
    # (PoolLossStep(pipelined=True) would also run the ingest kernels of batch i+1 under the compute of
    # batch i; measured slower here -- 0.47 against 0.39 ms per e2e step -- because the small kernels
    # delay CTAs of the persistent tensor-core kernel, and the e2e step is bound by the host link anyway)
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

    def timed_resident(steps, warmup):
        """value: inputs resident in HBM.  K graph replays queued back to back (the two buffer sets
        alternate, so a step never finds its inputs in L2), ONE CUDA-event pair around the region:
        the host runs ahead of the device, as it does inside a training loop."""
        sampler = ClockSampler(local_rank)
        sampler.start()
        for i in range(warmup):
            steps2[i % 2].run()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        last = None
        sampler.mark_begin()
        a.record()
        for i in range(steps):
            last = steps2[i % 2].run()
        b.record()
        barrier()
        sampler.mark_end()
        clocks = sampler.stop()
        return max_over_ranks(a.elapsed_time(b), "value_region_ms"), clocks, last

    def timed_e2e(steps, warmup):
        """e2e: every step copies its inputs from pinned host memory and reads the loss back.
        The H2D copy of step i+1 is enqueued (copy stream) before the host blocks on the loss of
        step i, so it overlaps step i's kernels; timed as one region over all K steps."""
        host_loss = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        read_ev = [torch.cuda.Event() for _ in range(2)]

        def loop(n):
            vals = []
            steps2[0].load(*host_args)
            for i in range(n):
                loss, _ = steps2[i % 2].run()                        # compute phase of batch i
                host_loss[i % 2].copy_(loss.reshape(1), non_blocking=True)   # device -> host read of its result
                read_ev[i % 2].record()
                if i + 1 < n:
                    steps2[(i + 1) % 2].load(*host_args)             # H2D copies + ingest of batch i+1
                if i >= 1:                                           # the host consumes loss i-1 while step i runs
                    read_ev[(i - 1) % 2].synchronize()
                    vals.append(float(host_loss[(i - 1) % 2]))
            read_ev[(n - 1) % 2].synchronize()
            vals.append(float(host_loss[(n - 1) % 2]))
            return vals
        loop(warmup)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        vals = loop(steps)
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b), "e2e_region_ms"), vals

    def timed_modular(steps, warmup, profile):
        for _ in range(warmup):
            step_modular()
        barrier()
        if profile:
            xr._native.lib().xr_fused_profile(1)
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

    # ---- device-resident throughput ---------------------------------------------------------------
    total_ms, clocks, last = timed_resident(args.steps, args.warmup)
    loss_val = float(last[0])
    m_a, m_rows = steps2[0].row_counts()
    c_cols = m_a + 1
    ms_per_step = total_ms / args.steps
    value = world * BATCH / (ms_per_step / 1e3)

    # ---- end to end: host buffers in, loss scalar out --------------------------------------------
    e2e_ms, e2e_vals = timed_e2e(args.steps, args.warmup)
    e2e_value = world * BATCH / (e2e_ms / args.steps / 1e3)
    h2d = sum(v.numel() * v.element_size() for v in host.values()) + host_tok.numel() * 2
    d2h = 4
    assert abs(e2e_vals[-1] - loss_val) <= 1e-6 * abs(loss_val), (e2e_vals[-1], loss_val)

    # ---- the trainer's whole compute_losses (trainer.py:213-264): train loss forward/backward AND
    #      LogitsStatistics + all seven losses, one sync-free graph replay per step ----------------
    mon = xr.PoolLossStep(emb, loss_fn, BATCH, SEQ_LEN, token_dtype=torch.bfloat16, monitor=True)
    mon.load(d_tok, d_idx["history_item_idx"], d_idx["pos_item_idx"], d_idx["neg_item_idx"])
    for _ in range(args.warmup):
        mon.run()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        flush.fill_(1)
        a.record()
        mon.run()
        b.record()
    barrier()
    mon_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev)) / args.steps
    mon_losses, mon_stats = mon.loss_dict()
    assert abs(float(mon_losses["loss/InfoNCELoss"]) - loss_val) <= 1e-4 * abs(loss_val), "monitor pass and train loss disagree"
    del mon

    # ---- the fused kernel alone (roofline leg) + the drop-in module path ------------------------
    # CUDA events recorded inside the library around every launch of the main fused kernel, on the
    # launching stream, over a timed loop of the SAME step issued without graph capture
    import ctypes

    mod_ms, mod_last = timed_modular(args.steps, args.warmup, profile=True)
    buf = (ctypes.c_float * 512)()
    n_prof = xr._native.lib().xr_fused_profile_read(buf, 512)
    xr._native.lib().xr_fused_profile(0)
    kern_ms = [buf[i] for i in range(max(n_prof, 0))]
    assert float(mod_last[0]) == loss_val, "graph-replayed step and module path disagree"

    flops = 4.0 * m_rows * c_cols * DIM          # scores + dQ, table frozen
    k_ms = statistics.mean(kern_ms) if kern_ms else float("nan")
    achieved = flops / (k_ms * 1e-3) / 1e12 if kern_ms else float("nan")
    line = {
        "metric": "train seq/sec (scoring-and-loss step)", "value": value, "unit": "seq/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": world * BATCH, "seq_len": SEQ_LEN,
                   "rows_M": m_rows, "candidates_C": c_cols,
                   "l2": "inputs larger than L2: two alternating buffer sets, ~2 x 125 MB touched per "
                         "pair of steps > 126 MB L2 (value and e2e); the kernel-only roofline leg "
                         "flushes L2 with a 256 MB write between launches",
                   "api": "PoolLossStep (xr_pool_step, CUDA-graph replay)",
                   "parallelism": f"dp{world} (independent batches, table replicated)"},
        "e2e": {"value": e2e_value, "unit": "seq/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                "note": "every step starts from pinned HOST inputs (H2D copy of batch i+1 on the copy stream "
                        "under the kernels of batch i); the loss of every step is copied to pinned host "
                        "memory and read by the host one step later (while the next step runs), as a "
                        "training loop logs it.  Bound by the host link: 20.3 MB per step at ~52 GB/s = "
                        "0.39 ms; a concurrent H2D stream also slows the tensor-core kernel (0.30 -> 0.37 ms "
                        "per step, measured), a D2D copy of the same size does not"},
        "gpu_launches": 7 * args.steps,   # xr_pool_step: compaction, plan, gather, diagonal, fused, finalize, row sum
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "fused_pool_kernel<InfoNCE> (tcgen05)",
                     "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved / peak_tf if kern_ms else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full
                     # capture of this kernel on this workload (profiles/ncu_fused_pool_kernel_r01c_raw.txt:
                     # 19.57 MB read + 6.43 MB written; algorithmic operand bytes 28.3 MB, partial dQ stays in L2)
                     "traffic": 26.0e6 if BATCH == 128 else None, "traffic_unit": "bytes",
                     "peak_source": f"{peak_src} sustained bf16 (MEASURED_PEAKS.json)",
                     "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms_per_step if kern_ms else None,
                     "algorithmic_flops_per_launch": flops},
        "loss": loss_val,
        "per_rank_region_ms": per_rank or None,
        "compute_losses": {"value": world * BATCH / (mon_ms / 1e3), "unit": "seq/s", "ms_per_step": mon_ms,
                           "gpu_launches_per_step": 20,
                           "note": "PoolLossStep(monitor=True): the step above PLUS LogitsStatistics and all "
                                   "seven losses (what trainer.py:250-263 logs every step) in the same graph "
                                   "replay; L2 flushed between steps",
                           "losses": {k: float(v) for k, v in mon_losses.items()}},
        "module_api": {"value": world * BATCH / (mod_ms / args.steps / 1e3), "unit": "seq/s",
                       "ms_per_step": mod_ms / args.steps,
                       "note": "same step through compute_embeds + InfoNCELoss + backward (the "
                               "reference's call sequence; one host sync for the row counts)"},
    }

    if not args.no_retrieval:
        line["retrieval"] = bench_retrieval(args, xr, dev, rank, world, peak_hbm)
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
        b0 = orc.synth_batch(N_ITEMS, BATCH, SEQ_LEN, dim=DIM, seed=0)
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

    if not args.no_retrieval:
        if rank == 0 and world == 1:   # exact cosine top-k on the host cores, bounded sample
            from oracle import cpu_baseline

            rows_s, q_s = min(args.catalog, 1_000_000), args.queries
            dt = cpu_baseline.time_exact_search(rows_s, q_s, 100, dim=DIM)
            line["retrieval"]["cpu_baseline"] = {
                "value": q_s / dt * rows_s / args.catalog, "unit": "queries/s", "cores": os.cpu_count(),
                "kind": "port",
                "sample": f"{q_s} queries x {rows_s} fp32 rows (chunked matmul + stable sort, torch CPU ops) in "
                          f"{dt:.2f} s = {q_s / dt:.1f} queries/s at that size, scaled linearly in the "
                          f"number of rows to the {args.catalog}-row catalog"}
    if rank == 0:
        print(json.dumps(line), flush=True)


def bench_retrieval(args, xr, dev, rank, world, peak_hbm):
    """Secondary metric of BASELINE.json: full-catalog top-100 queries/s, catalog sharded by rows
    across ranks, per-shard top-k merged after one NCCL all-gather (BASELINE configs[3])."""
    import torch

    from xfmr_rec_b200.dist import ShardedIndex, shard_range

    n, u, k = args.catalog, args.queries, 100
    lo, hi = shard_range(n, rank, world)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    shard = torch.randn((hi - lo, DIM), generator=g, device=dev, dtype=torch.float32).bfloat16()
    idx = xr.index.ExactIndex(xr.index.ExactIndexConfig(index_metric="cosine", dtype="bf16"), dev,
                              row_offset=lo)
    idx.catalog = shard  # rows are i.i.d. normal: norms ~ sqrt(384); normalise in place
    idx.catalog, _ = xr.ops.normalize_rows(shard, 1e-12, torch.bfloat16)
    del shard
    # sharded exchange over NVLink peer memory (all-gather + merge in one kernel); NCCL all-gather if
    # symmetric memory cannot be set up on this box
    sharded = ShardedIndex(idx, exchange="auto" if world > 1 else "nccl",
                           use_plan=os.environ.get("XR_BENCH_PLAN", "1") == "1")   # local search = one graph replay
    gq = torch.Generator(device=dev).manual_seed(99)
    q = torch.randn((u, DIM), generator=gq, device=dev)
    excl = None
    for _ in range(4):
        sharded.search_batch(q, excl, k)
    torch.cuda.synchronize()
    fused = xr.ops.score_groupmax_supported(q.bfloat16(), idx.catalog)
    if fused and not sharded.use_plan:   # (the event hook cannot be recorded inside a captured graph)
        xr._native.lib().xr_fused_profile(1)
    reps = 12
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    import gc

    gc.collect()
    gc.disable()        # a cyclic-GC pass of the host interpreter inside a 0.6 ms search shows up as a 5 ms search
    sharded.search_batch(q, excl, k)   # one more untimed search after the hooks above, then line the ranks up
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    for a, b in evs:
        a.record()
        s, i = sharded.search_batch(q, excl, k)
        b.record()
    torch.cuda.synchronize()
    gc.enable()
    in_order = [a.elapsed_time(b) for a, b in evs]
    per_rep = sorted(in_order)
    ms_mean = sum(per_rep) / reps
    ms = per_rep[reps // 2]   # median of 12 searches (the mean is reported beside it)
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    out = {"metric": "full-catalog top-100 queries/sec", "value": u / (ms / 1e3), "unit": "queries/s",
           "catalog_rows": n, "queries": u, "k": k, "ms_per_batch": ms, "ms_per_batch_mean": ms_mean,
           "ms_per_batch_min_max": [per_rep[0], per_rep[-1]], "timing": "median of 12 searches, CUDA events per search",
           "ms_per_search_in_order": [round(x, 4) for x in in_order],
           "dtype": "bf16",
           "path": "tcgen05 cta_group::2 group-max scoring + top groups re-scored + merge, local search as one CUDA-graph replay (+ exchange of (U,k) when sharded)"
           if fused else "scores (fp32-accumulate GEMM) + streaming top-k + merge",
           "catalog_bytes_per_rank": (hi - lo) * DIM * 2,
           "exchange": ("none (one GPU)" if world == 1 else
                        "NVLink peer memory: xr_topk_merge_peers after one device-side barrier"
                        if sharded._peer is not None else f"NCCL all-gather + merge ({sharded._peer_failed})")}
    if fused:
        import ctypes

        if sharded.use_plan:   # scoring-kernel time from a few eagerly issued local searches
            xr._native.lib().xr_fused_profile(1)
            for _ in range(4):
                idx.search_batch(q, None, k)
            torch.cuda.synchronize()
        buf = (ctypes.c_float * 512)()
        cnt = xr._native.lib().xr_fused_profile_read(buf, 512)
        xr._native.lib().xr_fused_profile(0)
        if cnt > 0:
            k_ms = sum(buf[i] for i in range(cnt)) / cnt
            flops = 2.0 * u * (hi - lo) * DIM
            byts = (hi - lo) * DIM * 2
            out["scoring_kernel"] = {
                "kernel": "score_gmax2_kernel (tcgen05 cta_group::2)" if u > 128 else "fused_pool_kernel<GMAX> (tcgen05)", "kernel_ms": k_ms,
                "TFLOP/s": flops / (k_ms * 1e-3) / 1e12,
                "catalog_GB/s": byts / (k_ms * 1e-3) / 1e9,
                "frac_of_measured_hbm": byts / (k_ms * 1e-3) / 1e9 / peak_hbm,
                "note": "U=256: 2*U*N*D FLOP vs N*D*2 catalog bytes read once; bound = max(tensor, HBM)"}
    return out


if __name__ == "__main__":
    main()
