#!/usr/bin/env python
"""bench.py -- headline benchmark of the pytorch_points hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Default workload = BASELINE.json configs[4] (`chamfer_b256_n8192`): Chamfer distance fwd+bwd, a
batch of 256 clouds of N=M=8192 points sharded over the GPUs (256/W clouds per rank: STRONG
scaling; at one GPU it is the largest single-GPU Chamfer configuration), metric = unique
point-pairs/s (B*N*M per step, whole job).  `--workload chamfer_b32_n2500` (configs[1], weak
scaling) and `chamfer_b32_n8192` (target shape) remain selectable and are reported, with their
end-to-end legs, under `extras`.  One "step" = nndistance forward (both directions + fused loss partial sums),
the backward scatter for loss = mean(dist1) + mean(dist2), and, when N>1, the exchange of the
two partial sums (peer-memory mailboxes; NCCL all-reduce with PP_LOSS_EXCHANGE=nccl).  Default:
the two-launch form pp_chamfer_fwd_bwd_uniform (backward folded into the index-resolving
kernel), checked against the four-launch sequence on the timed input first and timed next to
it (`fused_step` in the line); PP_FUSED_BWD=0 times the four-launch sequence only.

Legs printed in ONE JSON line by rank 0:
  value     device-resident: inputs already in HBM, C-ABI calls on the current stream;
  e2e       the reference-signature path a drop-in user runs: pinned HOST clouds copied in every
            step (double-buffered copy stream), the autograd Function behind `nndistance` /
            `sharded_chamfer_loss`, `loss.backward()`, and every step's loss read on the host
            (one step behind, like a logging training loop); the CUDA-graph pipeline
            (`pipeline.GraphedChamferStep`) and the plugin-level calls are reported next to it;
  roofline  dominant kernel (cs_rowpass_tc_kernel: tensor-core sweep; chamfer_fwd_kernel for small clouds),
            duration from CUDA events on the launching stream (library timing hooks); algorithmic 8 FLOP per
            unique pair against the FP32 pipe peak measured live, plus the TMEM-read and tensor-pipe figures;
  cpu_baseline  the CPU oracle port (oracle/pp_oracle.c, OpenMP) on the box's host cores;
  extras    (N=1 only) the other BASELINE.json configs: Chamfer B=32 N=M=8192, FPS
            16384->1024 + ball_query (B=16), group_knn k=16 (B=32 N=8192; B=4 N=131072).
`--impl reference` times the reference's CPU-side equivalent: the reference has NO CPU
implementation of this path, so the oracle port stands in (kind "port").
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (kind, per-GPU batch, N, M)
    "chamfer_b32_n2500": ("chamfer", 32, 2500, 2500),     # BASELINE.json configs[1]
    "chamfer_b32_n8192": ("chamfer", 32, 8192, 8192),     # north-star target shape
    "chamfer_b256_n8192": ("chamfer", 256, 8192, 8192),   # configs[4] at 1 GPU (batch is sharded /W)
}
L2_FLUSH_BYTES = 256 << 20


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_threads_setup():
    """All host threads for the OpenMP port, even under torch.distributed.run (which exports
    OMP_NUM_THREADS=1); must run before the oracle library is loaded."""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.setdefault("OMP_PROC_BIND", "spread")
    os.environ.setdefault("OMP_WAIT_POLICY", "active")
    import oracle
    lib = oracle.lib()
    lib.oracle_set_threads(n)  # libgomp may already have read OMP_NUM_THREADS=1
    return int(lib.oracle_get_max_threads())


def cpu_chamfer_step(a, b, total_batch):
    """One fwd+bwd pass of the oracle port over numpy clouds; returns the loss."""
    import numpy as np
    import oracle
    d1, d2, i1, i2 = oracle.chamfer_fwd(a, b)
    B, N = d1.shape
    M = d2.shape[1]
    gd1 = np.full((B, N), 1.0 / (total_batch * N), np.float32)
    gd2 = np.full((B, M), 1.0 / (total_batch * M), np.float32)
    oracle.chamfer_bwd(a, b, gd1, gd2, i1, i2)
    return float(d1.sum() / (total_batch * N) + d2.sum() / (total_batch * M))


def run_cpu_baseline(N, M, budget_s=12.0):
    """Oracle port on the host cores over a bounded sample of the workload."""
    import numpy as np
    from helpers import np32, uniform_cloud
    cpu_threads_setup()
    import oracle
    oracle.lib()
    a1, b1 = np32(uniform_cloud(1, N, 1001)), np32(uniform_cloud(1, M, 2001))
    cpu_chamfer_step(a1, b1, 1)  # warm up threads
    t0 = time.perf_counter()
    cpu_chamfer_step(a1, b1, 1)
    t1 = max(time.perf_counter() - t0, 1e-4)
    bs = int(max(1, min(32, budget_s / 3 / t1)))
    a, b = np32(uniform_cloud(bs, N, 1001)), np32(uniform_cloud(bs, M, 2001))
    best = None
    t_start = time.perf_counter()
    reps = 0
    while reps < 3 or time.perf_counter() - t_start < budget_s:
        t0 = time.perf_counter()
        cpu_chamfer_step(a, b, bs)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        reps += 1
    return {"value": bs * N * M / best, "unit": "point-pairs/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "oracle port (C + OpenMP, all host threads), Chamfer fwd+bwd on %d of the clouds, "
                      "N=M=%d, best of %d back-to-back runs over %.0f s (%.3f s each)" % (
                          bs, N, reps, time.perf_counter() - t_start, best)}


def reference_arm(args, world, rank):
    """--impl reference: the reference's CPU-side equivalent of the path on the host cores.  Rank 0
    alone runs it, with every host thread (the other ranks exit at once), on a bounded sample of the
    job's clouds; throughput in pairs/s does not depend on how many clouds the sample holds."""
    kind, B, N, M = WORKLOADS[args.workload]
    if rank != 0:
        return
    from helpers import np32, uniform_cloud
    cores = cpu_threads_setup()
    import oracle
    oracle.lib()
    a1, b1 = np32(uniform_cloud(1, N, 1001)), np32(uniform_cloud(1, M, 2001))
    t_warm = time.perf_counter()
    while time.perf_counter() - t_warm < 2.0:  # thread pool up, clocks settled
        cpu_chamfer_step(a1, b1, 1)
    t0 = time.perf_counter()
    cpu_chamfer_step(a1, b1, 1)
    t1 = max(time.perf_counter() - t0, 1e-4)
    budget = 120.0
    bs = int(max(1, min(B, budget / max(args.steps + args.warmup, 1) / t1)))
    a, b = np32(uniform_cloud(bs, N, 1001)), np32(uniform_cloud(bs, M, 2001))
    for _ in range(max(args.warmup, 1)):
        cpu_chamfer_step(a, b, bs)
    per_step = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        cpu_chamfer_step(a, b, bs)
        per_step.append(time.perf_counter() - t0)
    dt = sum(per_step) / max(len(per_step), 1)
    value = bs * N * M / dt
    scaling = "strong" if args.workload == "chamfer_b256_n8192" else "weak"
    sample = ("each step = oracle port (C + OpenMP, %d threads) Chamfer fwd+bwd over %d of the job's clouds, "
              "N=M=%d; the reference has no CPU implementation of this path (SURVEY.md D2)" % (cores, bs, N))
    line = {
        "impl": "reference", "metric": "chamfer_fwd_bwd_point_pairs_per_s", "value": value,
        "unit": "point-pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "best_ms_per_step": min(per_step) * 1e3 if per_step else None,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, B // world if scaling == "strong" else B, N, world),
                   "cpu_sample_clouds": bs},
        "cpu_baseline": {"value": value, "unit": "point-pairs/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "point-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(name, per_gpu_batch, N, world):
    """The `config.workload` string shared by both arms (the driver compares them)."""
    return "%s: Chamfer (nndistance) fwd+bwd, %d clouds in the job (%d per GPU), N=M=%d, uniform [0,1)^3, " \
           "loss = mean(dist1)+mean(dist2)" % (name, per_gpu_batch * world, per_gpu_batch, N)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="chamfer_b256_n8192", choices=sorted(WORKLOADS))
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        reference_arm(args, world, rank)
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: this benchmark has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    from helpers import uniform_cloud
    from pytorch_points_b200 import _C
    from pytorch_points_b200._ext import losses
    from pytorch_points_b200.dist import sharded_chamfer_loss

    kind, B, N, M = WORKLOADS[args.workload]
    if args.workload == "chamfer_b256_n8192":
        B = B // world           # configs[4]: the batch of 256 is sharded
        scaling = "strong"
    else:
        scaling = "weak"         # fixed per-GPU batch
    total_B = B * world
    pairs_per_step = float(total_B) * N * M

    # ---- inputs: seeded on CPU (identical bits on every path), one shard per rank
    a_host = uniform_cloud(B, N, 1001 + rank).pin_memory()
    b_host = uniform_cloud(B, M, 2001 + rank).pin_memory()
    a, b = a_host.to(dev), b_host.to(dev)
    d1 = torch.empty(B, N, device=dev); d2 = torch.empty(B, M, device=dev)
    i1 = torch.empty(B, N, dtype=torch.int32, device=dev); i2 = torch.empty(B, M, dtype=torch.int32, device=dev)
    sums = torch.zeros(2, device=dev)
    g1, g2 = torch.empty_like(a), torch.empty_like(b)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    gw = torch.tensor([1.0 / (total_B * N), 1.0 / (total_B * M)], device=dev)

    # The one collective: global [sum(dist1), sum(dist2)].  Default = the peer-memory exchange
    # (pp_loss_exchange_send / _wait: NVLink P2P stores into the peers' mailboxes); NCCL all-reduce
    # if the mailboxes cannot be mapped or PP_LOSS_EXCHANGE=nccl.
    exchange, exchange_note = None, "none (single GPU)"
    total = sums if world == 1 else torch.zeros(2, device=dev)
    if world > 1:
        exchange_note = "NCCL all-reduce of the 2 partial sums"
        if os.environ.get("PP_LOSS_EXCHANGE", "p2p") != "nccl":
            try:
                from pytorch_points_b200.dist import LossExchange
                exchange = LossExchange(dev)
                exchange_note = ("peer-memory exchange of the 2 partial sums (pp_loss_exchange_send/_wait: NVLink P2P "
                                 "stores into every rank's mailbox, summed in rank order; no collective call per step)")
            except Exception as e:  # noqa: BLE001
                exchange_note += " (peer exchange unavailable: %s)" % (repr(e)[:120],)

    # Fused step (default): forward + uniform backward in two launches
    # (pp_chamfer_fwd_bwd_uniform); PP_FUSED_BWD=0 = the four-launch sequence.  Checked against
    # the four-launch sequence on this very input before anything is timed.
    fused = os.environ.get("PP_FUSED_BWD", "1") != "0"
    fused_note = "disabled (PP_FUSED_BWD=0)"
    if fused:
        c1, c2 = torch.empty_like(a), torch.empty_like(b)
        losses.nmdistance_forward(a, b, d1, d2, i1, i2, sums=sums)
        losses.nmdistance_backward_uniform(a, b, c1, c2, gw, i1, i2)
        j1, j2 = i1.clone(), i2.clone()
        losses.nmdistance_forward_backward_uniform(a, b, d1, d2, i1, i2, sums, gw, g1, g2)
        err = max(float((g1 - c1).abs().max() / c1.abs().max()), float((g2 - c2).abs().max() / c2.abs().max()))
        if torch.equal(i1, j1) and torch.equal(i2, j2) and err <= 1e-5:
            fused_note = "on; gradients agree with the four-launch sequence to %.1e (rel. to max)" % err
        else:
            fused, fused_note = False, "CHECK FAILED (rel err %.3g): fell back to the four-launch sequence" % err
            sys.stderr.write("[bench] fused forward+backward disagrees with the separate kernels; not used\n")
        del c1, c2, j1, j2
        if world > 1:  # every rank must take the same path
            flag = torch.tensor([1 if fused else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if fused and int(flag.item()) == 0:
                fused, fused_note = False, "CHECK FAILED on another rank: four-launch sequence"

    def step_device_unfused():
        # forward (+ fused partial sums) -> exchange of the 2 sums -> backward with the two
        # constant upstream weights d(loss)/d(dist) = 1/(B_total*N), 1/(B_total*M)
        losses.nmdistance_forward(a, b, d1, d2, i1, i2, sums=sums)
        # the backward weights are constants, so the exchange overlaps the backward
        work = None
        if exchange is not None:
            exchange.send(sums)
        elif world > 1:
            total.copy_(sums)
            work = dist.all_reduce(total, async_op=True)
        losses.nmdistance_backward_uniform(a, b, g1, g2, gw, i1, i2)
        if exchange is not None:
            exchange.wait(total)
        elif work is not None:
            work.wait()

    def step_device_fused():
        # forward + finalize with the backward folded in (constant upstream weights), then the
        # exchange of the 2 sums
        losses.nmdistance_forward_backward_uniform(a, b, d1, d2, i1, i2, sums, gw, g1, g2)
        if exchange is not None:
            exchange.send(sums)
            exchange.wait(total)
        elif world > 1:
            total.copy_(sums)
            dist.all_reduce(total)

    step_device = step_device_fused if fused else step_device_unfused

    from pytorch_points_b200.pipeline import HostPrefetcher
    prefetcher = HostPrefetcher(dev, depth=2)
    prefetcher.prefetch((a_host, b_host))
    sums_host = torch.zeros(2).pin_memory()
    e_g1, e_g2 = torch.empty_like(a), torch.empty_like(b)

    def step_e2e():
        """Plugin-level step with HOST buffers: the reference-shaped `_ext.losses` calls (the drop-in
        boundary) fed from pinned host memory.  Inputs of THIS step were enqueued on the copy stream
        during the previous step; the next step's copies are enqueued now so that they overlap
        this step's kernels.  Ends with a device->host read of the step's result (the loss)."""
        xd, yd = prefetcher.get()
        losses.nmdistance_forward(xd, yd, d1, d2, i1, i2, sums=sums)
        work = None
        if exchange is not None:
            exchange.send(sums)
        elif world > 1:
            total.copy_(sums)
            work = dist.all_reduce(total, async_op=True)
        losses.nmdistance_backward_uniform(xd, yd, e_g1, e_g2, gw, i1, i2)
        prefetcher.release()
        prefetcher.prefetch((a_host, b_host))  # host work hidden behind the kernels just launched
        if exchange is not None:
            exchange.wait(total)
        elif work is not None:
            work.wait()
        sums_host.copy_(total, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(sums_host[0]) / (total_B * N) + float(sums_host[1]) / (total_B * M)

    from pytorch_points_b200.pipeline import GraphedChamferStep
    if world > 1:
        dist.all_reduce(sums)  # the communicator exists before the first replay
        torch.cuda.synchronize()
    graphed = GraphedChamferStep([(a_host, b_host)], total_batch=total_B, device=dev, world_size=world, exchange=exchange,
                                 fused_backward=fused)

    def step_e2e_graph():
        """The whole step (H2D x2 from pinned host, forward, backward, D2H of the loss sums) as one
        CUDA-graph replay followed by the host read of the loss."""
        return graphed.run()

    in_flight = []
    losses_read = [0]

    def step_e2e_pipelined():
        """Same graphs, software-pipelined two deep: enqueue step i+1 (its H2D copy was started during
        step i), then read step i's loss on the host while i+1 runs."""
        in_flight.append(graphed.submit())
        if len(in_flight) > 1:
            graphed.loss(in_flight.pop(0))
            losses_read[0] += 1

    def finish_pipelined():
        while in_flight:
            graphed.loss(in_flight.pop(0))
            losses_read[0] += 1

    def step_e2e_autograd_blocking():
        """The reference-signature step with nothing overlapped: host .to(device), autograd loss,
        backward, loss.item()."""
        x = a_host.to(dev, non_blocking=True).requires_grad_(True)
        y = b_host.to(dev, non_blocking=True).requires_grad_(True)
        loss = sharded_chamfer_loss(x, y, total_batch=total_B)
        loss.backward()
        return loss.item()

    ag_prefetcher = HostPrefetcher(dev, depth=2)
    ag_prefetcher.prefetch((a_host, b_host))
    from pytorch_points_b200.pipeline import HostScalarReader
    ag_pending = HostScalarReader(dev, depth=4)
    ag_read = [0]

    def step_e2e_autograd():
        """The reference-signature step as a training loop runs it: this step's clouds were copied from
        pinned host memory on the copy stream during the previous step (torch DataLoader-style
        prefetch), the autograd Function behind nndistance / sharded_chamfer_loss, loss.backward(),
        the next step's copies enqueued, then the PREVIOUS step's loss read on the host while this
        step runs (every step's loss is read inside the timed region: its device->host copy into pinned
        memory is enqueued right behind the forward -- pipeline.HostScalarReader -- because `.item()`
        would wait for everything launched so far and serialise host and GPU)."""
        xd, yd = ag_prefetcher.get()
        x, y = xd.detach().requires_grad_(True), yd.detach().requires_grad_(True)
        loss = sharded_chamfer_loss(x, y, total_batch=total_B, exchange=exchange)
        ag_pending.push(loss)
        loss.backward()
        ag_prefetcher.release()
        ag_prefetcher.prefetch((a_host, b_host))
        if len(ag_pending) > 1:
            ag_pending.pop()
            ag_read[0] += 1

    def finish_autograd():
        while len(ag_pending):
            ag_pending.pop()
            ag_read[0] += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, steps, warmup):
        for _ in range(warmup):
            step()
        barrier()
        evs = []
        for _ in range(steps):
            flush.zero_()  # evict L2 between timed iterations (inputs are smaller than L2)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            evs.append((e0, e1))
        barrier()
        total_ms = sum(x.elapsed_time(y) for x, y in evs)
        if world > 1:
            t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms / steps

    def timed_e2e(step, steps, warmup, finish=None):
        """One timed region over all K steps (every step's inputs arrive fresh over PCIe, so there is
        nothing to flush): CUDA events on the current stream, every step's loss is read on the host
        inside the region (`finish` collects the one still in flight of a pipelined step)."""
        for _ in range(warmup):
            step()
        if finish is not None:
            finish()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        if finish is not None:
            finish()
        e1.record()
        barrier()
        total_ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms / steps

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_step = timed(step_device, args.steps, args.warmup)
    ms_step_unfused = timed(step_device_unfused, args.steps, args.warmup) if fused else ms_step
    ms_e2e_plugin = timed_e2e(step_e2e, args.steps, args.warmup)
    ms_e2e_autograd_blocking = timed_e2e(step_e2e_autograd_blocking, min(args.steps, 50), 3)
    ag_read[0] = 0
    ms_e2e_autograd = timed_e2e(step_e2e_autograd, args.steps, args.warmup, finish=finish_autograd)
    assert ag_read[0] == args.steps + args.warmup, "every step's loss must be read on the host"
    if graphed is not None:
        ms_e2e_blocking = timed_e2e(step_e2e_graph, args.steps, args.warmup)
        losses_read[0] = 0
        ms_e2e_graph = timed_e2e(step_e2e_pipelined, args.steps, args.warmup, finish=finish_pipelined)
        assert losses_read[0] == args.steps + args.warmup, "every step's loss must be read on the host"
        e2e_api = ("pipeline.GraphedChamferStep.submit()/loss(): compute graph(s) (" +
                   ("pp_chamfer_fwd_bwd_uniform: forward + finalize with fused loss sums and backward" if fused else
                    "pp_chamfer_fwd + finalize with fused loss sums, pp_chamfer_bwd_uniform") +
                   ", [exchange of the sums when N>1,] D2H of the sums) on one "
                   "stream, copy graph (H2D of the NEXT step's two clouds from pinned host) on a second stream; "
                   "software-pipelined two deep: the host reads step i's loss while step i+1 runs -- every step copies "
                   "its inputs in and has its loss read on the host inside the timed region")
        loss_graph = step_e2e_graph()
    else:
        ms_e2e_graph, e2e_api, loss_graph = ms_e2e_plugin, "see plugin_api (CUDA-graph step is single-GPU only)", None
        ms_e2e_blocking = ms_e2e_plugin
    # separate short pass with the library's per-kernel CUDA events switched on, so the event
    # records do not perturb the two timed legs above
    _C.set_option("timing", 1)
    knames = ("chamfer_prep", "chamfer_fwd", "chamfer_finalize", "chamfer_bwd")
    for nm in knames:
        _C.timing_collect(nm)
    timed(step_device, min(args.steps, 50), 3)
    kt = {nm: _C.timing_collect(nm) for nm in knames}
    tensor_path = bool(_C.lib.pp_chamfer_last_path())
    # from ~3M points per call the library runs the uniform backward as its two streaming kernels again
    split_bwd = tensor_path and fused and B * (N + M) >= (3 << 20)
    _C.set_option("timing", 0)
    clocks = sampler.stop()
    loss_dev = float((total[0] / (total_B * N) + total[1] / (total_B * M)).item())
    loss_e2e = step_e2e()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel: chamfer_fwd_kernel, FP32 pipe bound
    n_all = args.steps + args.warmup
    fwd_ms = kt["chamfer_fwd"][0] / max(kt["chamfer_fwd"][1], 1)
    flops_per_launch = 8.0 * B * N * M            # 8 FLOP per unique pair
    ffma_ms, ffma_flop = _C.microbench(0, 4096, local_rank)
    mix_ms, mix_pairs = _C.microbench(3, 2048, local_rank)
    peak_tflops = ffma_flop / (ffma_ms * 1e-3) / 1e12
    pipe_pairs = mix_pairs / (mix_ms * 1e-3)
    achieved = flops_per_launch / (fwd_ms * 1e-3) / 1e12
    # DRAM bytes of one launch from the committed `ncu --set full` capture of this workload -- used only
    # if it was taken from the kernel source that is running now (hash of csrc/chamfer.cu + chamfer_sweep.cu + tc_common.cuh)
    traffic, traffic_note = None, "no ncu capture committed for this workload"
    try:
        import hashlib
        sha = hashlib.sha256(b"".join(open(os.path.join(ROOT, "pytorch_points_b200", "csrc", f), "rb").read()
                                      for f in ("chamfer.cu", "chamfer_sweep.cu", "tc_common.cuh"))).hexdigest()[:16]
        ent = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json"))).get(args.workload)
        if ent and world > 1:
            traffic_note = "the committed capture is of the single-GPU launch (256 clouds); not reported for a sharded batch"
        elif ent:
            if ent.get("chamfer_cu_sha16") == sha:
                traffic = ent.get("chamfer_fwd_dram_bytes_per_launch")
                traffic_note = "dram__bytes_read+write of %s, %s" % (ent.get("kernel", "?"), ent.get("capture", "?"))
            else:
                traffic_note = "stale capture (kernel source changed since %s): not reported" % ent.get("capture", "?")
    except Exception as e:  # noqa: BLE001
        traffic_note = "unreadable profiles/ncu_summary.json: %r" % (e,)
    per = lambda nm: kt[nm][0] / max(kt[nm][1], 1)
    others = {("chamfer_finalize(+fused backward)" if fused else "chamfer_finalize"): per("chamfer_finalize")}
    if not fused or split_bwd:
        others["chamfer_bwd(2 launches)"] = per("chamfer_bwd")
    if tensor_path:
        others["cs_prep_kernel"] = per("chamfer_prep")
    roofline = {
        "kernel": "cs_rowpass_tc_kernel" if tensor_path else "chamfer_fwd_kernel",
        "bound": "tensor" if tensor_path else "fp32",
        "achieved": achieved, "peak": peak_tflops,
        "unit": "TFLOP/s", "frac": achieved / peak_tflops, "traffic": traffic, "traffic_note": traffic_note,
        "peak_source": "FP32 FFMA peak measured live by pp_microbench (MEASURED_PEAKS.json holds no FP32 entry); achieved = "
                       "ALGORITHMIC work, 8 FLOP per unique pair (SURVEY.md 8d), per launch / CUDA-event duration",
        "kernel_ms": fwd_ms, "kernel_launches_timed": kt["chamfer_fwd"][1],
        "algorithmic_flop_per_launch": flops_per_launch,
        "pair_rate": B * N * M / (fwd_ms * 1e-3),
        "op_mix_ceiling_pairs_per_s": pipe_pairs,
        "frac_of_op_mix_ceiling": B * N * M / (fwd_ms * 1e-3) / pipe_pairs,
        "other_kernels_ms": others,
    }
    if tensor_path:
        # what bounds the tensor-core sweep: every accumulator element (one per pair and direction, 4 bytes) is read
        # out of TMEM once and goes through a minimum tree on the ALU pipe.  Both ceilings are measured live:
        # pp_microbench 7 = the kernel's tcgen05.ld shape alone, 8 = the loads + the 16 FMNMX3 per 32 values.
        ld_ms, ld_bytes = _C.microbench(7, 4000, local_rank)
        tree_ms, tree_bytes = _C.microbench(8, 4000, local_rank)
        tmem_bytes = 2.0 * 4.0 * B * N * M
        mma_flop = 2.0 * 2.0 * 16.0 * B * N * M  # two directions, K = 16 (3xTF32 split + norms, padded)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        roofline.update({
            "note": "distances come from tcgen05.mma kind::tf32 (3xTF32-split operands, K = 16, accumulators in TMEM); the FP32 "
                    "pipe only resolves the surviving candidates exactly, so the algorithmic 8-FLOP/pair rate is no longer "
                    "tied to the 0.667 ceiling of the exact FFMA chain.  The kernel's own ceiling is its epilogue: every "
                    "accumulator element is read from TMEM and reduced on the ALU pipe (FMNMX3 issues every 2 cycles); "
                    "epilogue_frac = achieved TMEM read rate / rate of a probe doing only those loads and that tree",
            "tmem_read_bytes_per_launch": tmem_bytes,
            "tmem_read_TBps": tmem_bytes / (fwd_ms * 1e-3) / 1e12,
            "tmem_ld_only_probe_TBps": ld_bytes / (ld_ms * 1e-3) / 1e12,
            "tmem_ld_plus_min_tree_probe_TBps": tree_bytes / (tree_ms * 1e-3) / 1e12,
            "epilogue_frac": (tmem_bytes / fwd_ms) / (tree_bytes / tree_ms),
            "tensor_flop_per_launch_executed": mma_flop,
            "tensor_TFLOPs_executed": mma_flop / (fwd_ms * 1e-3) / 1e12,
            "tensor_frac_of_measured_bf16_peak": (mma_flop / (fwd_ms * 1e-3) / 1e12 / peaks["bf16_tflops"]) if peaks.get("bf16_tflops") else None,
            "tensor_peak_note": "kind::tf32 runs at half the bf16 rate; MEASURED_PEAKS.json holds the bf16 figure only",
        })
    else:
        roofline["note"] = ("the reference rounding order needs 6 FP32 lane-ops per pair (3 sub, 1 mul, 2 fma = 8 FLOP in "
                            "12 FLOP slots), so 0.667 of the FFMA peak is the hard ceiling; op_mix_ceiling is that bound "
                            "measured live with the packed FADD2/FMUL2/FFMA2+FMNMX3 mix")

    line = {
        "metric": "chamfer_fwd_bwd_point_pairs_per_s", "value": pairs_per_step / (ms_step * 1e-3),
        "unit": "point-pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, B, N, world),
                   "loss_exchange": exchange_note if world > 1 else "none (single GPU)",
                   "global_batch": total_B,
                   "l2": "flushed between timed steps (256 MiB write); inputs %.1f MB per GPU" % (
                       (a_host.numel() + b_host.numel()) * 4 / 1e6),
                   "timing": "per-step CUDA events on the current stream, max over ranks"},
        "e2e": {"value": pairs_per_step / (ms_e2e_autograd * 1e-3), "unit": "point-pairs/s", "ms_per_step": ms_e2e_autograd,
                "h2d_bytes_per_step": world * (a_host.numel() + b_host.numel()) * 4, "d2h_bytes_per_step": world * 4,
                "api": "reference-signature path: pipeline.HostPrefetcher (pinned host -> device on a copy stream, double "
                       "buffered) + dist.sharded_chamfer_loss (the torch.autograd Function behind nndistance, loss sums "
                       "exchanged inside the autograd node when N>1: NVLink peer mailboxes, torch.distributed.all_reduce if peer "
                       "mapping is unavailable) + loss.backward(); each loss travels to pinned host memory right behind its "
                       "forward (pipeline.HostScalarReader) and is read while step i+1 runs -- every step copies its "
                       "inputs in and has its loss read inside the timed region",
                "autograd_blocking_api": {"value": pairs_per_step / (ms_e2e_autograd_blocking * 1e-3), "ms_per_step": ms_e2e_autograd_blocking,
                                          "api": "host .to(device) + dist.sharded_chamfer_loss + loss.backward() + loss.item(), nothing overlapped"},
                "graph_api": {"value": pairs_per_step / (ms_e2e_graph * 1e-3), "ms_per_step": ms_e2e_graph,
                              "h2d_bytes_per_step": world * (a_host.numel() + b_host.numel()) * 4, "d2h_bytes_per_step": world * 8,
                              "api": e2e_api},
                "graph_blocking_api": {"value": pairs_per_step / (ms_e2e_blocking * 1e-3), "ms_per_step": ms_e2e_blocking,
                                       "api": "pipeline.GraphedChamferStep.run(): same graphs, the host waits for each step's loss "
                                              "before enqueueing the next step"},
                "plugin_api": {"value": pairs_per_step / (ms_e2e_plugin * 1e-3), "ms_per_step": ms_e2e_plugin,
                               "api": "pipeline.HostPrefetcher (pinned host -> device, double buffered) + _ext.losses.nmdistance_forward / "
                                      "nmdistance_backward_uniform (the reference-shaped plugin boundary) + D2H of the loss sums"},
                "timing": "one CUDA-event region over all K steps; every step copies its inputs from pinned host memory and has its loss read on the host"},
        "fused_step": {"state": fused_note, "ms_per_step_four_launch_sequence": ms_step_unfused},
        "gpu_launches": (((3 if fused and not split_bwd else 5) if tensor_path else (2 if fused else 4)) + (2 if exchange is not None else 0)) * args.steps,
        "gpu_launches_note": ("per step: " + (("cs_prep_kernel, cs_rowpass_tc_kernel, cs_finalize_kernel" +
                                               ("<fused backward>" if fused else ""))
                                              if tensor_path else
                                              ("chamfer_fwd_kernel, chamfer_finalize_kernel" + ("<fused backward>" if fused else "")))
                              + ("" if fused and not split_bwd else ", chamfer_bwd_kernel<0>, <1>")
                              + (", lx_send_kernel, lx_wait_kernel" if exchange is not None else "")),
        "clocks": clocks, "roofline": roofline,
        "loss": {"device_leg": loss_dev, "e2e_plugin_leg": loss_e2e, "e2e_graph_leg": loss_graph},
    }

    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = run_cpu_baseline(N, M)
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"error": repr(e)}
    if world == 1 and not args.no_extras:
        try:
            line["extras"] = run_extras(dev, _C, peak_tflops, pipe_pairs)
        except Exception as e:  # noqa: BLE001
            line["extras"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_extras(dev, _C, peak_tflops, pipe_pairs):
    """The remaining BASELINE.json configs and north-star shapes on one GPU (kernel-level)."""
    import torch
    from helpers import uniform_cloud
    from pytorch_points_b200._ext import losses, sampling
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def timeit(fn, iters=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    out = {}
    # Chamfer fwd+bwd at configs[1] (AtlasNet shape) and at the north-star target shape, device-resident
    # and end to end (same two e2e paths as the headline workload)
    from pytorch_points_b200.dist import sharded_chamfer_loss
    from pytorch_points_b200.pipeline import GraphedChamferStep, HostPrefetcher, HostScalarReader
    for (B, N) in [(32, 2500), (32, 8192)]:
        a_host, b_host = uniform_cloud(B, N, 1).pin_memory(), uniform_cloud(B, N, 2).pin_memory()
        a, b = a_host.to(dev), b_host.to(dev)
        d1 = torch.empty(B, N, device=dev); d2 = torch.empty(B, N, device=dev)
        i1 = torch.empty(B, N, dtype=torch.int32, device=dev); i2 = torch.empty(B, N, dtype=torch.int32, device=dev)
        gw = torch.full((2,), 1.0 / (B * N), device=dev)
        g1, g2 = torch.empty_like(a), torch.empty_like(b)
        sums = torch.zeros(2, device=dev)

        def chamfer_step():
            if os.environ.get("PP_FUSED_BWD", "1") != "0":
                losses.nmdistance_forward_backward_uniform(a, b, d1, d2, i1, i2, sums, gw, g1, g2)
            else:
                losses.nmdistance_forward(a, b, d1, d2, i1, i2, sums=sums)
                losses.nmdistance_backward_uniform(a, b, g1, g2, gw, i1, i2)
        ms = timeit(chamfer_step)  # the step as it runs (programmatic dependent launch between its kernels)
        # the forward kernel alone: a separate pass with the library's CUDA events (launches fully serialised there)
        _C.set_option("timing", 1)
        _C.timing_collect("chamfer_fwd")
        timeit(chamfer_step, iters=10, warm=0)
        tot, cnt = _C.timing_collect("chamfer_fwd")
        _C.set_option("timing", 0)
        kms = tot / max(cnt, 1)
        # e2e, K steps in one region, every step's inputs from pinned host memory, every loss read on the host
        K, pf, pending = 40, HostPrefetcher(dev, depth=2), HostScalarReader(dev, depth=4)
        pf.prefetch((a_host, b_host))

        def ag_step():
            xd, yd = pf.get()
            x, y = xd.detach().requires_grad_(True), yd.detach().requires_grad_(True)
            loss = sharded_chamfer_loss(x, y, total_batch=B)
            pending.push(loss)  # D2H of the loss right behind the forward; read one step later
            loss.backward()
            pf.release()
            pf.prefetch((a_host, b_host))
            if len(pending) > 1:
                pending.pop()
        graphed = GraphedChamferStep([(a_host, b_host)], total_batch=B, device=dev, world_size=1, exchange=None,
                                     fused_backward=os.environ.get("PP_FUSED_BWD", "1") != "0")
        inflight = []

        def graph_step():
            inflight.append(graphed.submit())
            if len(inflight) > 1:
                graphed.loss(inflight.pop(0))

        def region(step, drain):
            for _ in range(5):
                step()
            drain()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(K):
                step()
            drain()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / K

        def drain_ag():
            while len(pending):
                pending.pop()

        def drain_graph():
            while inflight:
                graphed.loss(inflight.pop(0))
        ms_ag, ms_gr = region(ag_step, drain_ag), region(graph_step, drain_graph)
        out["chamfer_fwd_bwd_B%d_N%d" % (B, N)] = {
            "ms_per_step": ms, "point_pairs_per_s": B * N * N / (ms * 1e-3), "fwd_kernel_ms": kms,
            "fwd_kernel_tflops": 8.0 * B * N * N / (kms * 1e-3) / 1e12,
            "fwd_kernel_frac_of_fp32_peak": 8.0 * B * N * N / (kms * 1e-3) / 1e12 / peak_tflops,
            "fwd_kernel_frac_of_op_mix_ceiling": B * N * N / (kms * 1e-3) / pipe_pairs,
            "e2e_autograd_ms_per_step": ms_ag, "e2e_autograd_point_pairs_per_s": B * N * N / (ms_ag * 1e-3),
            "e2e_graph_ms_per_step": ms_gr, "e2e_graph_point_pairs_per_s": B * N * N / (ms_gr * 1e-3),
            "h2d_bytes_per_step": (a_host.numel() + b_host.numel()) * 4,
            "e2e_note": "autograd = HostPrefetcher + sharded_chamfer_loss + backward, loss read one step behind; "
                        "graph = pipeline.GraphedChamferStep.submit()/loss(); both copy the clouds from pinned host "
                        "memory every step and read every step's loss on the host"}
        del a, b, d1, d2, i1, i2, g1, g2, graphed, pf

    # config 3: FPS 16384 -> 1024 (+gather) and ball_query r=0.2 nsample=32, B=16
    B, N, m = 16, 16384, 1024
    x = uniform_cloud(B, N, 3).to(dev)
    idx = torch.empty(B, m, dtype=torch.int32, device=dev)
    temp = torch.empty(B, N, device=dev)

    def fps_step():
        temp.fill_(1e10)
        sampling.furthest_sampling(m, 0, x, temp, idx)
    _C.set_option("timing", 1)
    _C.timing_collect("fps")
    ms = timeit(fps_step, iters=5, warm=2)
    tot, cnt = _C.timing_collect("fps")
    _C.set_option("timing", 0)
    kms = tot / max(cnt, 1)
    alg_bytes = 20.0 * N * (m - 1) * B
    smem_ms, smem_bytes = _C.microbench(4, 2048, dev.index)  # shared-memory read bandwidth, all 148 SMs
    smem_per_sm = smem_bytes / (smem_ms * 1e-3) / 1e9 / 148.0
    cluster, per_thread = _C.fps_last_plan()  # what the library actually launched
    sms_used = B * max(cluster, 1)
    out["fps_B16_N16384_m1024"] = {
        "ms_per_step": ms, "samples_per_s": B * m / (ms * 1e-3), "kernel_ms": kms,
        "us_per_round": kms * 1e3 / (m - 1),
        "algorithmic_GBps": alg_bytes / (kms * 1e-3) / 1e9,
        "smem_GBps_per_sm_measured": smem_per_sm, "sms_holding_the_clouds": sms_used,
        "cluster_width": cluster, "points_per_thread": per_thread,
        "frac_of_smem_roofline": alg_bytes / (kms * 1e-3) / 1e9 / (smem_per_sm * sms_used),
        "note": "cloud and running minima are register-resident across a thread-block cluster: no per-round "
                "memory traffic; algorithmic bytes = 20*N per selected sample (SURVEY.md 8d)"}
    ctr = torch.gather(x, 1, idx.long().unsqueeze(-1).expand(B, m, 3)).contiguous()
    ms = timeit(lambda: sampling.ball_query(ctr, x, 0.2, 32))
    out["ball_query_B16_N16384_M1024_r0.2_ns32"] = {"ms_per_step": ms, "centres_per_s": B * m / (ms * 1e-3),
                                                     "pair_tests_upper_bound_per_s": B * m * N / (ms * 1e-3)}
    feats = x.transpose(1, 2).contiguous()
    gout = torch.empty(B, 3, m, device=dev)
    ms = timeit(lambda: sampling.gather_forward(B, 3, N, m, feats, idx, gout))
    out["gather_B16_C3_m1024"] = {"ms_per_step": ms}
    # the whole sampling + grouping work of a set-abstraction level:
    # FPS with fused gather, then ONE fused ball-query + grouping kernel, C = 64 feature channels
    from pytorch_points_b200 import network as ppn
    f64 = uniform_cloud(B, N, 5, c=64).transpose(1, 2).contiguous().to(dev)
    grouper, grouper_ops = ppn.QueryAndGroup(0.2, 32), ppn.QueryAndGroup(0.2, 32, fused=False)

    def sa_fused():
        ctr_ = ppn.furthest_point_sample(x, m, NCHW=False)[1]
        return grouper(x, ctr_, f64)

    def sa_ops():
        i_ = ppn.FurthestPointSampling.apply(x, m, 0)
        ctr_ = ppn.gather_points(feats, i_).transpose(1, 2).contiguous()
        return grouper_ops(x, ctr_, f64)
    ms_f, ms_o = timeit(sa_fused, iters=5, warm=2), timeit(sa_ops, iters=5, warm=2)
    ms_g = timeit(lambda: grouper(x, ctr, f64))
    ms_go = timeit(lambda: grouper_ops(x, ctr, f64))
    # point-major staging: the features copied once per level to (B,N,C), every scale gathers full lines
    ms_stage = timeit(lambda: ppn.stage_features(f64))
    f64_pm = ppn.stage_features(f64)
    ms_gpm = timeit(lambda: grouper(x, ctr, f64, features_pm=f64_pm))
    # the calls above are host bound at this size (a Python autograd Function + a 140 MB allocation per call):
    # the kernels themselves, from the library's CUDA events
    _C.set_option("timing", 1)
    kern = {}
    for name_, fn_ in (("query_and_group_kernel_ms", lambda: grouper(x, ctr, f64)),
                       ("query_and_group_staged_kernel_ms", lambda: grouper(x, ctr, f64, features_pm=f64_pm))):
        _C.timing_collect("query_group")
        for _ in range(5):
            fn_()
        torch.cuda.synchronize()
        tot_, cnt_ = _C.timing_collect("query_group")
        kern[name_] = tot_ / max(cnt_, 1)
    _C.timing_collect("channels_to_points")
    for _ in range(5):
        ppn.stage_features(f64)
    torch.cuda.synchronize()
    tot_, cnt_ = _C.timing_collect("channels_to_points")
    kern["stage_features_kernel_ms"] = tot_ / max(cnt_, 1)
    _C.set_option("timing", 0)
    out_bytes = 4.0 * B * (3 + 64) * m * 32
    out["sa_stage_B16_N16384_m1024_r0.2_ns32_C64"] = {
        "fused_ms": ms_f, "op_by_op_ms": ms_o, "query_and_group_fused_ms": ms_g,
        "query_and_group_op_by_op_ms": ms_go, "query_and_group_output_GBps": out_bytes / (ms_g * 1e-3) / 1e9,
        "stage_features_ms": ms_stage, "query_and_group_staged_ms": ms_gpm,
        "query_and_group_staged_output_GBps": out_bytes / (ms_gpm * 1e-3) / 1e9,
        **kern,
        "query_and_group_staged_kernel_output_GBps": out_bytes / (max(kern["query_and_group_staged_kernel_ms"], 1e-9) * 1e-3) / 1e9,
        "note": "op_by_op = the reference's kernel sequence (FPS, gather, ball_query, 2x group_points, "
                "subtract, cat) on this repo's single kernels"}
    # feature propagation: three_nn of all 16384 points against the 1024
    # sampled centres, then three_interpolate of C = 64 coarse features back onto the 16384 points
    d3 = torch.empty(B, N, 3, device=dev); i3 = torch.empty(B, N, 3, dtype=torch.int32, device=dev)
    ms_nn = timeit(lambda: sampling.three_nn_wrapper(B, N, m, x, ctr, d3, i3))
    w3 = torch.rand(B, N, 3, device=dev)
    w3 = (w3 / w3.sum(-1, keepdim=True)).contiguous()
    coarse = uniform_cloud(B, m, 6, c=64).transpose(1, 2).contiguous().to(dev)
    interp = torch.empty(B, 64, N, device=dev)
    ms_it = timeit(lambda: sampling.three_interpolate_wrapper(B, 64, m, N, coarse, i3, w3, interp))
    out["fp_stage_B16_n16384_m1024_C64"] = {
        "three_nn_ms": ms_nn, "three_nn_pairs_per_s": float(B) * N * m / (ms_nn * 1e-3),
        "three_interpolate_ms": ms_it,
        "three_interpolate_GBps": (4.0 * B * 64 * N + 24.0 * B * N) / (ms_it * 1e-3) / 1e9}
    del x, ctr, feats, f64

    # group_knn k=16: target shape and config 4
    for (B, N, iters) in [(32, 8192, 5), (4, 131072, 2)]:
        p = uniform_cloud(B, N, 4).to(dev)
        _C.set_option("timing", 1)
        _C.timing_collect("knn")
        ms = timeit(lambda: sampling.knn(16, p, p), iters=iters, warm=1)
        tot, cnt = _C.timing_collect("knn")
        _C.set_option("timing", 0)
        kms = tot / max(cnt, 1)
        _C.set_option("knn_stats", 1)
        sampling.knn(16, p, p)
        visited, total = _C.knn_stats()
        _C.set_option("knn_stats", 0)
        _C.set_option("knn_prune", 0)
        ms_dense = timeit(lambda: sampling.knn(16, p, p), iters=2, warm=1)
        _C.set_option("knn_prune", 1)
        # the tensor-core path on the same input (knn_tc.cu; automatic only for 16 < k <= 32), with its stage times
        _C.set_option("knn_tc", 1)
        try:
            ms_tc = timeit(lambda: sampling.knn(16, p, p), iters=iters, warm=1)
            _C.set_option("timing", 1)
            for n_ in ("knn_sort", "knn_prep", "knn_seed", "knn", "knn_select"):
                _C.timing_collect(n_)
            sampling.knn(16, p, p)
            torch.cuda.synchronize()
            stages = {n_: _C.timing_collect(n_)[0] for n_ in ("knn_sort", "knn_prep", "knn_seed", "knn", "knn_select")}
            _C.set_option("timing", 0)
            _C.set_option("knn_stats", 1)
            sampling.knn(16, p, p)
            vis_tc, tot_tc = _C.knn_stats()
        finally:
            _C.set_option("knn_stats", 0)
            _C.set_option("timing", 0)
            _C.set_option("knn_tc", -1)
        tc = {"ms_per_step": ms_tc, "stage_ms": stages, "blocks_128x128_evaluated_frac": vis_tc / max(tot_tc, 1.0),
              "evaluated_pairs_per_s_in_flagging_kernel": float(B) * N * N * vis_tc / max(tot_tc, 1.0) / max(stages["knn"] * 1e-3, 1e-9),
              "note": "tcgen05.mma flagging pass over Morton-ordered 128x128 blocks + exact seed / resolution kernels "
                      "(knn_tc = 1); same bits as the ordered sweep"}
        # where the tensor path is the automatic choice: lists wider than 16 -- K = k + 1 = 17 is what the
        # snapshot's DenseEdgeConv asks for (network/layers.py:52 with k = 16)
        for kk in ((17, 32) if (B, N) == (32, 8192) else (17,)):
            _C.set_option("knn_tc", 0)
            ms_sweep = timeit(lambda: sampling.knn(kk, p, p), iters=3, warm=1)
            _C.set_option("knn_tc", -1)
            ms_auto = timeit(lambda: sampling.knn(kk, p, p), iters=3, warm=1)
            tc["k%d_ms_per_step_automatic_choice" % kk] = ms_auto
            tc["k%d_ms_per_step_ordered_sweep" % kk] = ms_sweep
        out["knn_k16_B%d_N%d" % (B, N)] = {
            "ms_per_step": ms, "point_pairs_per_s": float(B) * N * N / (ms * 1e-3), "sweep_kernel_ms": kms,
            "algorithmic_tflops": 8.0 * B * N * N / (ms * 1e-3) / 1e12,
            "algorithmic_frac_of_fp32_peak": 8.0 * B * N * N / (ms * 1e-3) / 1e12 / peak_tflops,
            "tiles_evaluated_frac": visited / max(total, 1.0),
            "evaluated_pairs_per_s_in_sweep_kernel": float(B) * N * N * visited / max(total, 1.0) / (kms * 1e-3),
            "ms_per_step_without_pruning": ms_dense,
            "tensor_core_path": tc,
            "note": "algorithmic pairs = B*M*N (SURVEY.md 8d); exact bounding-box pruning skips the tiles that "
                    "cannot hold a neighbour, so the algorithmic rate may exceed the FP32 pipe"}
        del p
    try:
        out["reference_cuda_kernels"] = run_reference_cuda(dev, timeit)
    except Exception as e:  # noqa: BLE001
        out["reference_cuda_kernels"] = {"unavailable": repr(e)[:200]}
    return out


def run_reference_cuda(dev, timeit):
    """Baseline only: the reference's OWN CUDA kernels (unmodified sources compiled for sm_100a by
    oracle/build_ref.sh into oracle/_ref/, prebuilt -- /root/reference is not read here) timed on
    the same shapes, kernels only.  Shows what "recompile the reference for B200" gives."""
    import torch
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "ref_losses.so")):
        return {"unavailable": "oracle/_ref not built"}
    sys.path.insert(0, ref_dir)
    import ref_losses
    import ref_sampling
    from helpers import uniform_cloud
    res = {}
    for (B, N) in [(32, 2500), (32, 8192), (256, 8192)]:
        a, b = uniform_cloud(B, N, 1).to(dev), uniform_cloud(B, N, 2).to(dev)
        d1 = torch.zeros(B, N, device=dev); d2 = torch.zeros(B, N, device=dev)
        i1 = torch.zeros(B, N, dtype=torch.int32, device=dev); i2 = torch.zeros(B, N, dtype=torch.int32, device=dev)
        gd = torch.full((B, N), 1.0 / (B * N), device=dev)
        g1, g2 = torch.zeros_like(a), torch.zeros_like(b)

        def step():
            ref_losses.nmdistance_forward(a, b, d1, d2, i1, i2)
            ref_losses.nmdistance_backward(a, b, g1, g2, gd, gd, i1, i2)
        ms = timeit(step, iters=5, warm=2)
        res["chamfer_fwd_bwd_B%d_N%d" % (B, N)] = {"ms_per_step": ms, "point_pairs_per_s": float(B) * N * N / (ms * 1e-3)}
    x = uniform_cloud(16, 16384, 3).to(dev)
    idx = torch.empty(16, 1024, dtype=torch.int32, device=dev)
    temp = torch.empty(16, 16384, device=dev)

    def fps_step():
        temp.fill_(1e10)
        ref_sampling.furthest_sampling(1024, 0, x, temp, idx)
    ms = timeit(fps_step, iters=3, warm=1)
    res["fps_B16_N16384_m1024"] = {"ms_per_step": ms, "samples_per_s": 16 * 1024 / (ms * 1e-3)}
    ctr = torch.gather(x, 1, idx.long().unsqueeze(-1).expand(16, 1024, 3)).contiguous()
    ms = timeit(lambda: ref_sampling.ball_query(ctr, x, 0.2, 32), iters=5, warm=2)
    res["ball_query_B16_N16384_M1024_r0.2_ns32"] = {"ms_per_step": ms}
    # the reference's QueryAndGroup sequence on its own kernels (ball_query, 2x group_points, subtract, cat), C = 64
    f64 = uniform_cloud(16, 16384, 5, c=64).transpose(1, 2).contiguous().to(dev)
    xt = x.transpose(1, 2).contiguous()

    def ref_query_and_group():
        i_ = ref_sampling.ball_query(ctr, x, 0.2, 32)
        g_ = ref_sampling.group_points(xt, i_)
        g_ -= ctr.transpose(1, 2).unsqueeze(-1)
        return torch.cat([g_, ref_sampling.group_points(f64, i_)], dim=1)
    ms = timeit(ref_query_and_group, iters=5, warm=2)
    res["query_and_group_B16_N16384_m1024_r0.2_ns32_C64"] = {"ms_per_step": ms}
    # feature propagation: three_nn + three_interpolate
    d3 = torch.empty(16, 16384, 3, device=dev); i3 = torch.empty(16, 16384, 3, dtype=torch.int32, device=dev)
    ms = timeit(lambda: ref_sampling.three_nn_wrapper(16, 16384, 1024, x, ctr, d3, i3), iters=5, warm=2)
    res["three_nn_B16_n16384_m1024"] = {"ms_per_step": ms, "pairs_per_s": 16.0 * 16384 * 1024 / (ms * 1e-3)}
    w3 = torch.rand(16, 16384, 3, device=dev)
    coarse = uniform_cloud(16, 1024, 6, c=64).transpose(1, 2).contiguous().to(dev)
    interp = torch.empty(16, 64, 16384, device=dev)
    ms = timeit(lambda: ref_sampling.three_interpolate_wrapper(16, 64, 1024, 16384, coarse, i3, w3, interp), iters=5, warm=2)
    res["three_interpolate_B16_C64_n16384_m1024"] = {"ms_per_step": ms}
    return res


if __name__ == "__main__":
    sys.exit(main())
