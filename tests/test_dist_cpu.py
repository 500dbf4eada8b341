"""N>1 host logic on CPU: two gloo ranks shard a batch, run the (oracle-backed) local Chamfer op,
all-reduce the partial sums and back-propagate.  Checks that the loss equals the single-process
global mean and that every rank's gradients are its slice of the global gradient."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class _OracleNN(torch.autograd.Function):
    """CPU stand-in with the signature of nndistance, backed by the oracle (test-only)."""

    @staticmethod
    def forward(ctx, a, b):
        import oracle
        d1, d2, i1, i2 = oracle.chamfer_fwd(a.detach().numpy(), b.detach().numpy())
        ctx.save_for_backward(a, b, torch.from_numpy(i1), torch.from_numpy(i2))
        return torch.from_numpy(d1), torch.from_numpy(d2), torch.from_numpy(i1), torch.from_numpy(i2)

    @staticmethod
    def backward(ctx, g1, g2, _a, _b):
        import oracle
        a, b, i1, i2 = ctx.saved_tensors
        x, y = oracle.chamfer_bwd(a.numpy(), b.numpy(), g1.contiguous().numpy(), g2.contiguous().numpy(),
                                  i1.numpy(), i2.numpy())
        return torch.from_numpy(x), torch.from_numpy(y)


def _global_reference(a, b):
    a = a.clone().requires_grad_(True)
    b = b.clone().requires_grad_(True)
    d1, d2, _, _ = _OracleNN.apply(a, b)
    loss = d1.mean() + d2.mean()
    loss.backward()
    return loss.item(), a.grad, b.grad


def _worker(rank, world, port, B, N, M, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import uniform_cloud
    from pytorch_points_b200.dist import shard_batch, shard_range, sharded_chamfer_loss
    a, b = uniform_cloud(B, N, 1), uniform_cloud(B, M, 2)
    la = shard_batch(a).clone().requires_grad_(True)
    lb = shard_batch(b).clone().requires_grad_(True)
    loss = sharded_chamfer_loss(la, lb, local_op=_OracleNN.apply)   # total_batch discovered by all-reduce
    loss.backward()
    lo, hi = shard_range(B, rank, world)
    ref_loss, ga, gb = _global_reference(a, b)
    ok = abs(loss.item() - ref_loss) <= 1e-6 * abs(ref_loss)
    ok = ok and torch.allclose(la.grad, ga[lo:hi], rtol=1e-5, atol=1e-9)
    ok = ok and torch.allclose(lb.grad, gb[lo:hi], rtol=1e-5, atol=1e-9)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_shard_range_covers_batch():
    from pytorch_points_b200.dist import shard_range
    for B in (1, 5, 8, 32, 255):
        for W in (1, 2, 3, 8):
            spans = [shard_range(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_chamfer_two_ranks_gloo():
    import oracle
    oracle.lib()  # build once, before forking
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, 5, 64, 48, out), nprocs=world, join=True)
    assert all(out.get(r) is True for r in range(world)), dict(out)


def test_single_process_falls_back_to_no_collective():
    from helpers import uniform_cloud
    from pytorch_points_b200.dist import sharded_chamfer_loss
    a, b = uniform_cloud(3, 32, 3).requires_grad_(True), uniform_cloud(3, 40, 4).requires_grad_(True)
    loss = sharded_chamfer_loss(a, b, local_op=_OracleNN.apply)
    ref, ga, gb = _global_reference(a.detach(), b.detach())
    loss.backward()
    assert abs(loss.item() - ref) < 1e-7 and torch.allclose(a.grad, ga) and torch.allclose(b.grad, gb)
