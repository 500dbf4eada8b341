"""Host -> device input pipeline: double-buffered, copy-stream prefetch of pinned host batches.

The hot path itself never touches host memory; this helper is what a training loop puts in front
of it so that the PCIe copy of step i+1's clouds overlaps step i's kernels (bench.py's `e2e` leg
uses it).  Buffers are reused, so no allocation happens in steady state."""
import torch

from . import _C


def _indexed(device):
    """torch.device with an explicit ordinal (the library's entry points take one)."""
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


class HostPrefetcher:
    """Cycles through `depth` sets of device buffers.  `prefetch(tensors)` enqueues the H2D copies
    of a tuple of pinned host tensors on a private stream; `get()` makes the current stream wait
    for the oldest outstanding set and returns its device tensors."""

    def __init__(self, device, depth=2):
        self.device = _indexed(device)
        self.depth = depth
        self.stream = torch.cuda.Stream(self.device)
        self._bufs = [None] * depth
        # one event pair per slot, re-recorded every cycle (creating events costs more than recording them)
        self._ready = [torch.cuda.Event() for _ in range(depth)]
        self._consumed = [torch.cuda.Event() for _ in range(depth)]
        self._used = [False] * depth
        self._head = 0   # next slot to fill
        self._tail = 0   # next slot to hand out
        self._inflight = 0

    def prefetch(self, host_tensors):
        if self._inflight >= self.depth:
            raise RuntimeError("HostPrefetcher: all %d buffer sets are in flight" % self.depth)
        slot = self._head
        if self._bufs[slot] is None or any(b.shape != h.shape or b.dtype != h.dtype
                                           for b, h in zip(self._bufs[slot], host_tensors)):
            self._bufs[slot] = tuple(torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host_tensors)
        if self._used[slot]:
            self.stream.wait_event(self._consumed[slot])  # previous user of this slot is done
        handle = self.stream.cuda_stream
        for b, h in zip(self._bufs[slot], host_tensors):
            if h.is_contiguous() and h.device.type == "cpu":
                # plain cudaMemcpyAsync on the copy stream through the library: ~3 us of host time,
                # tensor.copy_ under a stream context costs 10-20
                _C.check(_C.lib.pp_memcpy_async(_C.ptr(b), _C.ptr(h), h.numel() * h.element_size(), self.device.index,
                                                _C._vp(handle)), "pp_memcpy_async")
            else:
                with torch.cuda.stream(self.stream):
                    b.copy_(h, non_blocking=True)
        self._ready[slot].record(self.stream)
        self._head = (slot + 1) % self.depth
        self._inflight += 1

    def get(self):
        if self._inflight == 0:
            raise RuntimeError("HostPrefetcher: nothing was prefetched")
        slot = self._tail
        torch.cuda.current_stream(self.device).wait_event(self._ready[slot])
        self._tail = (slot + 1) % self.depth
        self._inflight -= 1
        self._last = slot
        return self._bufs[slot]

    def release(self):
        """Call after the last kernel that reads the tensors returned by the latest `get()` has been
        enqueued: records the event the copy stream waits on before overwriting that slot."""
        self._consumed[self._last].record(torch.cuda.current_stream(self.device))
        self._used[self._last] = True


class HostScalarReader:
    """Reads small device results (a loss) on the host WITHOUT draining the stream.

    `tensor.item()` enqueues its copy behind everything already launched on the stream and waits for all
    of it: a loop that reads step i-1's loss after launching step i therefore runs the GPU and the host
    strictly one after the other.  `push(t)` instead copies `t` into a pinned slot right where it is
    produced in stream order (call it straight after the forward, before `backward()`), followed by an
    event; `pop()` waits for the oldest slot's event only -- long complete by the time the next step has
    been enqueued -- and returns the value(s) as Python floats."""

    def __init__(self, device, depth=4, numel=1):
        self.device = _indexed(device)
        self._host = [torch.zeros(numel, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self._events = [torch.cuda.Event() for _ in range(depth)]
        self._head = self._tail = self._inflight = 0
        self.depth, self.numel = depth, numel

    def __len__(self):
        return self._inflight

    def push(self, t):
        if self._inflight >= self.depth:
            raise RuntimeError("HostScalarReader: all %d slots are in flight" % self.depth)
        slot = self._head
        if t.numel() != self.numel or t.dtype != torch.float32 or not t.is_contiguous():
            t = t.detach().reshape(-1).to(torch.float32).contiguous()
        stream = torch.cuda.current_stream(self.device)
        _C.check(_C.lib.pp_memcpy_async(_C.ptr(self._host[slot]), _C.ptr(t), 4 * self.numel, self.device.index,
                                        _C._vp(stream.cuda_stream)), "pp_memcpy_async")
        self._events[slot].record(stream)
        self._head = (slot + 1) % self.depth
        self._inflight += 1

    def pop(self):
        if self._inflight == 0:
            raise RuntimeError("HostScalarReader: nothing was pushed")
        slot = self._tail
        self._events[slot].synchronize()
        self._tail = (slot + 1) % self.depth
        self._inflight -= 1
        h = self._host[slot]
        return float(h[0]) if self.numel == 1 else [float(v) for v in h]


class GraphedChamferStep:
    """One Chamfer training step -- H2D of both clouds from pinned host memory, nndistance forward
    with the fused loss sums, the backward scatter for loss = mean(dist1) + mean(dist2), and the
    D2H read of the loss -- captured in CUDA graphs and replayed with two launches per step.

    The kernels are launch-latency sized at AtlasNet shapes (B=32, N=M=2500: ~0.1 ms of GPU work
    in four kernels), so a Python loop around them is host-bound; the graphs remove the per-call
    host work.  Two buffer sets alternate: while the compute graph of step i runs, the copy graph
    of step i+1 moves the next clouds over PCIe on a second stream.  `host_pairs` is one or two
    (xyz1, xyz2) pairs of PINNED host tensors (two = the loader fills one while the other is in
    flight); results of the latest step: `self.grad1`, `self.grad2`, `self.dist1` ... (valid in
    `compute_stream` order).

    Two ways to drive it: `run()` = one step, blocks until its loss is on the host; or
    `t = submit()` / `loss(t)` = software pipeline of depth two -- enqueue step i+1, then read step
    i's loss while i+1 runs (each step still copies its inputs in and its loss out).
    """

    def __init__(self, host_pairs, total_batch=None, device=None, group=None, world_size=1, exchange=None,
                 fused_backward=True):
        """With `world_size` > 1 every rank builds its own step over its shard of the batch.  With an
        `exchange` (dist.LossExchange) the global loss sums travel through peer memory and the whole
        step -- forward, send, backward, wait, D2H -- is ONE graph.  Without it the 8-byte NCCL
        all-reduce is NOT captured: the step is split into a forward graph and a backward graph
        and the all-reduce is issued eagerly between the two replays (asynchronously, so it
        overlaps the backward, whose weights are constants).
        `fused_backward` (default): the forward and the uniform backward are the two launches of
        `nmdistance_forward_backward_uniform`; False = the four-launch sequence forward, finalize,
        backward x2."""
        from ._ext import losses
        self.world_size, self.group, self.exchange = world_size, group, exchange
        dev = torch.device(device if device is not None else torch.cuda.current_device())
        if dev.type != "cuda":
            raise RuntimeError("GraphedChamferStep needs a CUDA device")
        if isinstance(host_pairs[0], torch.Tensor):
            host_pairs = [tuple(host_pairs)]
        host_pairs = [tuple(p) for p in host_pairs]
        if len(host_pairs) == 1:
            host_pairs = host_pairs * 2
        for a, b in host_pairs:
            if not (a.is_pinned() and b.is_pinned()):
                raise RuntimeError("GraphedChamferStep: host tensors must be pinned")
        self.host_pairs = host_pairs
        B, N, _ = host_pairs[0][0].shape
        M = host_pairs[0][1].shape[1]
        tb = total_batch if total_batch is not None else B
        self.scale = (1.0 / (tb * N), 1.0 / (tb * M))
        self.device = dev
        self.xyz1 = [torch.empty(B, N, 3, device=dev) for _ in range(2)]
        self.xyz2 = [torch.empty(B, M, 3, device=dev) for _ in range(2)]
        self.dist1 = torch.empty(B, N, device=dev)
        self.dist2 = torch.empty(B, M, device=dev)
        self.idx1 = torch.empty(B, N, dtype=torch.int32, device=dev)
        self.idx2 = torch.empty(B, M, dtype=torch.int32, device=dev)
        self.sums = torch.zeros(2, device=dev)
        self.gw = torch.tensor(self.scale, device=dev)
        self.grad1 = torch.empty(B, N, 3, device=dev)
        self.grad2 = torch.empty(B, M, 3, device=dev)
        self.sums_host = [torch.zeros(2).pin_memory() for _ in range(2)]  # per buffer set
        # The step owns its scratch: the graphs bake its address in, nothing else ever writes it, and
        # every forward leaves the packed keys all-ones again -> filled once, PP_CHAMFER_WS_CLEAN after.
        from . import _C
        self.workspace = torch.full((max(int(_C.lib.pp_chamfer_fwd_workspace_bytes(B, N, M)), 16),), 0xFF,
                                    dtype=torch.uint8, device=dev)
        self.compute_stream = torch.cuda.Stream(dev)
        self.copy_stream = torch.cuda.Stream(dev)
        # chamfer_fwd, chamfer_finalize[, chamfer_bwd<0>, chamfer_bwd<1>] (+ lx_send, lx_wait)
        self.fused_backward = bool(fused_backward)
        self.launches = (2 if self.fused_backward else 4) + (2 if exchange is not None else 0)

        def copy_body(s):
            self.xyz1[s].copy_(self.host_pairs[s][0], non_blocking=True)
            self.xyz2[s].copy_(self.host_pairs[s][1], non_blocking=True)

        def fwd_body(s):
            if self.fused_backward:
                losses.nmdistance_forward_backward_uniform(self.xyz1[s], self.xyz2[s], self.dist1, self.dist2,
                                                           self.idx1, self.idx2, self.sums, self.gw, self.grad1,
                                                           self.grad2, workspace=self.workspace, workspace_clean=True)
            else:
                losses.nmdistance_forward(self.xyz1[s], self.xyz2[s], self.dist1, self.dist2, self.idx1, self.idx2,
                                          sums=self.sums, workspace=self.workspace, workspace_clean=True)

        def bwd_body(s):
            if not self.fused_backward:
                losses.nmdistance_backward_uniform(self.xyz1[s], self.xyz2[s], self.grad1, self.grad2, self.gw,
                                                   self.idx1, self.idx2)

        self.total = torch.zeros(2, device=dev) if exchange is not None else self.sums

        def compute_body(s):
            fwd_body(s)
            if exchange is not None:
                exchange.send(self.sums)
            bwd_body(s)
            if exchange is not None:
                exchange.wait(self.total)
            self.sums_host[s].copy_(self.total, non_blocking=True)

        cur = torch.cuda.current_stream(dev)
        self.copy_stream.wait_stream(cur)
        self.compute_stream.wait_stream(cur)
        with torch.cuda.stream(self.copy_stream):
            copy_body(0)
            copy_body(1)
        self.copy_stream.synchronize()
        if exchange is not None:
            import torch.distributed as dist
            dist.barrier(group=group)  # ranks enter the first exchange together (its poll is bounded)
        with torch.cuda.stream(self.compute_stream):
            for _ in range(2):  # warm-up on the capture stream (also creates its key workspace)
                compute_body(0)
        self.compute_stream.synchronize()
        self.copy_graph, self.compute_graph = [], []
        for s in range(2):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.copy_stream):
                copy_body(s)
            self.copy_graph.append(g)
            if world_size == 1 or exchange is not None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.compute_stream):
                    compute_body(s)
                self.compute_graph.append(g)
            else:
                gf = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gf, stream=self.compute_stream):
                    fwd_body(s)
                gb = None  # fused: the backward already ran inside the forward graph
                if not self.fused_backward:
                    gb = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gb, stream=self.compute_stream):
                        bwd_body(s)
                self.compute_graph.append((gf, gb))
        self.copied = [torch.cuda.Event(), torch.cuda.Event()]
        self.done = [torch.cuda.Event(), torch.cuda.Event()]
        self.slot = 0
        self._primed = False

    def _launch_copy(self, s):
        with torch.cuda.stream(self.copy_stream):
            # set s was last read by the compute enqueued two submits ago (no-op before that)
            self.copy_stream.wait_event(self.done[s])
            self.copy_graph[s].replay()
            self.copied[s].record(self.copy_stream)

    def submit(self):
        """Enqueue one step on the current buffer set and start moving the next set's clouds.
        Returns a ticket for `loss()`; does not block."""
        s = self.slot
        if not self._primed:
            self._launch_copy(s)
            self._primed = True
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(self.copied[s])
            if self.world_size == 1 or self.exchange is not None:
                self.compute_graph[s].replay()
            else:
                import torch.distributed as dist
                gf, gb = self.compute_graph[s]
                gf.replay()
                work = dist.all_reduce(self.sums, group=self.group, async_op=True)
                if gb is not None:
                    gb.replay()
                work.wait()
                self.sums_host[s].copy_(self.sums, non_blocking=True)
            self.done[s].record(self.compute_stream)
        self._launch_copy(1 - s)
        self.slot = 1 - s
        return s

    def loss(self, ticket):
        """Wait for the step behind `ticket` and return its loss (host float).  A ticket is valid
        until the second `submit()` after it."""
        self.done[ticket].synchronize()
        h = self.sums_host[ticket]
        value = float(h[0]) * self.scale[0] + float(h[1]) * self.scale[1]
        if value != value and self.exchange is not None and self.exchange.timed_out():
            raise RuntimeError("GraphedChamferStep: the peer-memory loss exchange gave up waiting for a rank "
                               "(library option lx_timeout_ms); the loss of this step is undefined")
        return value

    def run(self):
        """One step, blocking: returns the loss (host float) once it is on the host."""
        return self.loss(self.submit())
