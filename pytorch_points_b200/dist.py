"""Batch-sharded execution over the GPUs of one box (one process per GPU, torch.distributed).

The reference has no distributed code at all.  Every op on the hot path is
independent per batch element, so rank r of W simply owns a contiguous slice of the batch and
runs the kernels on it; outputs stay sharded.  The ONLY collective is the all-reduce (SUM) of
the two Chamfer partial sums [sum(dist1), sum(dist2)] -- 8 bytes over NVLink -- needed for the
global loss mean.  Backward needs no collective: d(loss)/d(dist) = 1/(B_total*N) is a constant.
"""
import torch
import torch.distributed as dist


def shard_range(total_batch, rank, world_size):
    """Contiguous split of the batch dimension; the first (total % world) ranks get one extra."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size %r/%r" % (rank, world_size))
    base, rem = divmod(total_batch, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(tensor, rank=None, world_size=None):
    """The slice of `tensor` (batch-first) this rank owns."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(tensor.shape[0], rank, world_size)
    return tensor[lo:hi]


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group)
    return 1


def sharded_chamfer_loss(xyz1, xyz2, total_batch=None, group=None, local_op=None, exchange=None):
    """Mean Chamfer loss  mean(dist1) + mean(dist2)  over a batch sharded across ranks.

    xyz1 (b_local, N, 3), xyz2 (b_local, M, 3): this rank's clouds.  `total_batch` is the global
    batch size (defaults to the all-reduced sum of local batch sizes).  Returns a scalar that is
    identical on every rank; its gradient w.r.t. the local clouds is the local share of the global
    mean, so `loss.backward()` needs no communication.  `local_op` defaults to the CUDA
    `nndistance`; tests inject a CPU stand-in to exercise the host logic under gloo.  `exchange`: a
    `LossExchange` of the same ranks; the two partial sums then travel through NVLink peer memory instead of
    a `torch.distributed.all_reduce`."""
    world = _world(group)
    if total_batch is None:
        tb = torch.tensor([float(xyz1.shape[0])], device=xyz1.device)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.SUM, group=group)
        total_batch = int(tb.item())
    n, m = xyz1.shape[1], xyz2.shape[1]
    w1, w2 = 1.0 / (total_batch * max(n, 1)), 1.0 / (total_batch * max(m, 1))
    if local_op is None:
        # one library call: forward, fused loss sums and -- the weights being constants -- the backward;
        # the 8-byte all-reduce of the sums happens inside the same autograd node
        from .network.model_loss import chamfer_weighted_loss
        how = None if world == 1 else (exchange if exchange is not None else (group if group is not None else True))
        return chamfer_weighted_loss(xyz1, xyz2, w1, w2, how)[0]
    d1, d2, _, _ = local_op(xyz1, xyz2)
    sums = torch.stack([d1.sum(), d2.sum()])
    local = sums[0] * w1 + sums[1] * w2
    if world == 1:
        return local
    total = sums.detach().clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    global_loss = total[0] * w1 + total[1] * w2
    # value = global mean, gradient = this rank's share of it
    return local + (global_loss - local).detach()


def allreduce_chamfer_sums(sums, group=None):
    """In-place all-reduce (SUM) of the fused [sum(dist1), sum(dist2)] buffer written by
    `pp_chamfer_fwd`; the device-resident path bench.py measures."""
    if _world(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


class LossExchange:
    """All-reduce (SUM) of the 2-float Chamfer partial sums over NVLink peer memory
    (csrc/loss_exchange.cu): every rank stores its sums straight into every peer's mailbox and
    polls its own.  Two tiny graph-capturable kernels per step, no collective-library call.

        lx = LossExchange(device)            # collective: every rank of `group` must call it
        ... pp_chamfer_fwd(..., sums) ...
        lx.send(sums)                        # right after the forward
        ... backward kernels ...
        lx.wait(total)                       # total (2 floats, device) = sum over ranks

    Raises at construction when peer mapping is impossible (no CUDA IPC / no P2P); callers then
    keep using `torch.distributed.all_reduce`."""

    def __init__(self, device, group=None):
        import ctypes
        from . import _C
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("LossExchange needs an initialised torch.distributed process group")
        self._C = _C
        self.device = torch.device(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nbytes = _C.lib.pp_loss_exchange_handle_bytes()
        handle = (ctypes.c_ubyte * nbytes)()
        box = ctypes.c_void_p()
        err = None
        try:
            _C.check(_C.lib.pp_loss_exchange_create(ctypes.byref(box), handle, self.device.index), "pp_loss_exchange_create")
        except RuntimeError as e:  # still take part in the collectives below
            err = repr(e)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (bytes(handle), err), group=group)
        self.mailbox = box
        self.peers = (ctypes.c_void_p * self.world)()
        if not any(g[1] for g in gathered):
            for p, (hb, _) in enumerate(gathered):
                if p == self.rank:
                    self.peers[p] = box.value
                    continue
                ptr = ctypes.c_void_p()
                buf = (ctypes.c_ubyte * nbytes).from_buffer_copy(hb)
                try:
                    _C.check(_C.lib.pp_loss_exchange_open(buf, ctypes.byref(ptr), self.device.index), "pp_loss_exchange_open")
                    self.peers[p] = ptr.value
                except RuntimeError as e:
                    err = repr(e)
                    break
        else:
            err = err or "a peer could not create its mailbox"
        # agree on the outcome so that either every rank uses the exchange or none does
        flags = [None] * self.world
        dist.all_gather_object(flags, err, group=group)
        bad = [f for f in flags if f]
        if bad:
            raise RuntimeError("LossExchange unavailable: " + bad[0])
        _C.check(_C.lib.pp_loss_exchange_bind(self.mailbox, self.peers, self.world, self.device.index), "pp_loss_exchange_bind")
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        dist.barrier(group=group)  # every mailbox is bound before anyone sends

    def send(self, sums):
        C = self._C
        C.check(C.lib.pp_loss_exchange_send(C.ptr(sums), self.mailbox, self.rank, self.world, self.device.index,
                                            C.stream_of(self.device)), "pp_loss_exchange_send")

    def wait(self, out):
        C = self._C
        C.check(C.lib.pp_loss_exchange_wait(self.mailbox, self.world, C.ptr(out), C.ptr(self.status), self.device.index,
                                            C.stream_of(self.device)), "pp_loss_exchange_wait")
        return out

    def timed_out(self):
        """True if any wait() so far gave up on a peer (synchronises).  The poll is bounded by the
        library option "lx_timeout_ms" (default 10 minutes, 0 = unbounded); a wait that gives up
        writes NaN sums and sets this sticky flag."""
        return bool(self.status.item())

    def close(self):
        """Unmap the peers' mailboxes and free this rank's (synchronises the device).  Every rank
        should call it once no step is in flight any more; safe to call twice."""
        if getattr(self, "mailbox", None) is not None and self.mailbox.value:
            C = self._C
            C.lib.pp_loss_exchange_close(self.mailbox, self.peers, self.rank, self.world, self.device.index)
            self.mailbox = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 -- interpreter shutdown: the driver reclaims the memory anyway
            pass
