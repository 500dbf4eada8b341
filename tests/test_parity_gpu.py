"""GPU parity tests: the CUDA kernels (called through the reference-shaped Python layer, i.e.
through the C ABI) against the CPU oracle on identical seeded inputs.

Bar (BASELINE.json north_star): indices bit-exact, distances bit-exact (same rounding order),
gradients within 1e-5 relative error (atomic summation order differs, as in the reference)."""
import os

import numpy as np
import pytest
import torch

from helpers import lattice_cloud, np32, sphere_cloud, uniform_cloud, with_duplicates

pytestmark = pytest.mark.gpu

GRAD_RTOL = 1e-5  # north_star tolerance for fp32 gradients


@pytest.fixture(scope="module")
def pp():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pytorch_points_b200  # noqa: F401  raises if libpp_b200.so is missing
    from pytorch_points_b200 import network
    return network


def dev(t):
    return t.cuda()


def assert_grad_close(got, want, what):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    scale = max(np.abs(want).max(), 1e-30)
    err = np.abs(got - want).max() / scale
    assert err <= GRAD_RTOL, "%s: max rel err %.3g" % (what, err)


# --------------------------------------------------------------------------- chamfer
CHAMFER_CASES = [
    # (B, N, M, maker, seed)
    (1, 1, 1, uniform_cloud, 1),
    (2, 7, 5, uniform_cloud, 2),
    (3, 33, 1, uniform_cloud, 3),
    (2, 1, 129, uniform_cloud, 4),
    (2, 128, 128, uniform_cloud, 5),
    (2, 513, 700, uniform_cloud, 6),
    (4, 2500, 2500, uniform_cloud, 7),       # config 2 shape, smaller batch
    (2, 1000, 3000, sphere_cloud, 8),
    (2, 2048, 2048, lattice_cloud, 9),       # massive exact ties
    (1, 4097, 4099, uniform_cloud, 10),
    (1, 20000, 17001, uniform_cloud, 11),    # one big odd-sized cloud pair: many query splits per reference block
    (70, 300, 40, sphere_cloud, 12),         # many tiny clouds
]


@pytest.mark.parametrize("B,N,M,maker,seed", CHAMFER_CASES)
def test_chamfer_forward_bit_exact(pp, oracle_mod, B, N, M, maker, seed):
    a, b = maker(B, N, 1000 + seed), maker(B, M, 2000 + seed)
    d1, d2, i1, i2 = pp.nndistance(dev(a), dev(b))
    e1, e2, j1, j2 = oracle_mod.chamfer_fwd(np32(a), np32(b))
    assert i1.dtype == torch.int32 and i2.dtype == torch.int32
    assert np.array_equal(np32(i1), j1), "idx1"
    assert np.array_equal(np32(i2), j2), "idx2"
    assert np.array_equal(np32(d1).view(np.uint32), e1.view(np.uint32)), "dist1 bits"
    assert np.array_equal(np32(d2).view(np.uint32), e2.view(np.uint32)), "dist2 bits"


_VARIANTS = [1, 2, 5, 13, 14, 21, 22, 25, 31, 32, 35, 51]


@pytest.mark.parametrize("variant", _VARIANTS)
def test_chamfer_forward_all_variants(pp, oracle_mod, variant):
    from pytorch_points_b200 import _C
    a = with_duplicates(uniform_cloud(2, 1500, 11))
    b = with_duplicates(uniform_cloud(2, 1100, 12))
    b[:, :100] = a[:, :100]  # exact zero distances too
    e1, e2, j1, j2 = oracle_mod.chamfer_fwd(np32(a), np32(b))
    _C.set_option("chamfer_variant", variant)
    try:
        d1, d2, i1, i2 = pp.nndistance(dev(a), dev(b))
    finally:
        _C.set_option("chamfer_variant", 0)
    assert np.array_equal(np32(i1), j1) and np.array_equal(np32(i2), j2)
    assert np.array_equal(np32(d1), e1) and np.array_equal(np32(d2), e2)


def test_chamfer_identical_clouds(pp, oracle_mod):
    a = uniform_cloud(2, 777, 13)
    d1, d2, i1, i2 = pp.nndistance(dev(a), dev(a.clone()))
    assert (np32(d1) == 0).all() and (np32(d2) == 0).all()
    ar = np.arange(777, dtype=np.int32)[None].repeat(2, 0)
    assert np.array_equal(np32(i1), ar) and np.array_equal(np32(i2), ar)


@pytest.mark.parametrize("c", [1, 2, 5])
def test_chamfer_generic_point_dim(pp, oracle_mod, c):
    a, b = uniform_cloud(2, 300, 14, c=c), uniform_cloud(2, 600, 15, c=c)
    d1, d2, i1, i2 = pp.nndistance(dev(a), dev(b))
    e1, e2, j1, j2 = oracle_mod.chamfer_fwd(np32(a), np32(b))
    assert np.array_equal(np32(i1), j1) and np.array_equal(np32(i2), j2)
    assert np.array_equal(np32(d1), e1) and np.array_equal(np32(d2), e2)


def test_chamfer_generic_kernel_matches_fast_path(pp):
    from pytorch_points_b200 import _C
    a, b = dev(lattice_cloud(2, 900, 16)), dev(lattice_cloud(2, 1300, 17))
    fast = pp.nndistance(a, b)
    _C.set_option("chamfer_generic", 1)
    try:
        slow = pp.nndistance(a, b)
    finally:
        _C.set_option("chamfer_generic", 0)
    for x, y in zip(fast, slow):
        assert torch.equal(x, y)


def test_chamfer_empty_inputs(pp):
    a = torch.zeros(2, 0, 3).cuda()
    b = uniform_cloud(2, 10, 18).cuda()
    d1, d2, i1, i2 = pp.nndistance(a, b)
    assert d1.shape == (2, 0) and d2.shape == (2, 10)
    assert (d2 == 0).all() and (i2 == 0).all()  # reference leaves the Python-side zeros in place
    z = pp.nndistance(torch.zeros(0, 5, 3).cuda(), torch.zeros(0, 4, 3).cuda())
    assert z[0].shape == (0, 5)


def test_chamfer_noncontiguous_inputs(pp, oracle_mod):
    a = uniform_cloud(2, 400, 19).transpose(1, 2).contiguous().transpose(1, 2)  # (B,N,3) non-contiguous view
    b = uniform_cloud(2, 500, 20)[:, ::2]                                      # strided
    assert not a.is_contiguous() and not b.is_contiguous()
    d1, d2, i1, i2 = pp.nndistance(dev(a), dev(b))
    e1, e2, j1, j2 = oracle_mod.chamfer_fwd(np32(a), np32(b))
    assert np.array_equal(np32(i1), j1) and np.array_equal(np32(i2), j2)


@pytest.mark.parametrize("B,N,M,maker", [(2, 300, 200, uniform_cloud), (3, 2500, 2500, uniform_cloud),
                                          (2, 1024, 1024, lattice_cloud)])
def test_chamfer_backward(pp, oracle_mod, B, N, M, maker):
    a, b = maker(B, N, 21), maker(B, M, 22)
    ga, gb = dev(a).requires_grad_(True), dev(b).requires_grad_(True)
    d1, d2, i1, i2 = pp.nndistance(ga, gb)
    w1, w2 = uniform_cloud(B, N, 23, c=1)[..., 0], uniform_cloud(B, M, 24, c=1)[..., 0]
    ((d1 * dev(w1)).sum() + (d2 * dev(w2)).sum()).backward()
    e1, e2 = oracle_mod.chamfer_bwd(np32(a), np32(b), np32(w1), np32(w2), np32(i1), np32(i2))
    assert_grad_close(np32(ga.grad), e1, "gradxyz1")
    assert_grad_close(np32(gb.grad), e2, "gradxyz2")


def test_chamfer_loss_mean_backward_matches_torch_autograd(pp):
    """End-to-end AtlasNet-style loss against plain PyTorch autograd on the same device."""
    a, b = dev(uniform_cloud(2, 500, 25)).requires_grad_(True), dev(uniform_cloud(2, 400, 26)).requires_grad_(True)
    d1, d2, _, _ = pp.nndistance(a, b)
    (d1.mean() + d2.mean()).backward()
    a2, b2 = a.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    D = ((a2[:, :, None, :] - b2[:, None, :, :]) ** 2).sum(-1)
    (D.min(2)[0].mean() + D.min(1)[0].mean()).backward()
    assert_grad_close(np32(a.grad), np32(a2.grad), "grad a")
    assert_grad_close(np32(b.grad), np32(b2.grad), "grad b")


def test_fused_mean_loss_matches_nndistance_path(pp):
    """chamfer_mean_loss / chamfer_sums (fused reduction + uniform backward) vs the reference-shaped
    nndistance path followed by .mean()."""
    a0, b0 = uniform_cloud(3, 700, 61), uniform_cloud(3, 500, 62)
    a, b = dev(a0).requires_grad_(True), dev(b0).requires_grad_(True)
    d1, d2, _, _ = pp.nndistance(a, b)
    l1 = d1.mean() + d2.mean()
    l1.backward()
    a2, b2 = dev(a0).requires_grad_(True), dev(b0).requires_grad_(True)
    l2 = pp.chamfer_mean_loss(a2, b2)
    (3.0 * l2).backward()
    assert abs(l1.item() - l2.item()) <= 1e-6 * abs(l1.item())
    assert_grad_close(np32(a2.grad), 3.0 * np32(a.grad), "fused grad a")
    assert_grad_close(np32(b2.grad), 3.0 * np32(b.grad), "fused grad b")


def test_sharded_loss_single_rank_on_gpu(pp):
    from pytorch_points_b200.dist import sharded_chamfer_loss
    a0, b0 = uniform_cloud(2, 300, 63), uniform_cloud(2, 400, 64)
    a, b = dev(a0).requires_grad_(True), dev(b0).requires_grad_(True)
    loss = sharded_chamfer_loss(a, b)
    loss.backward()
    a2, b2 = dev(a0).requires_grad_(True), dev(b0).requires_grad_(True)
    d1, d2, _, _ = pp.nndistance(a2, b2)
    ref = d1.mean() + d2.mean()
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-6 * abs(ref.item())
    assert_grad_close(np32(a.grad), np32(a2.grad), "sharded grad a")
    assert_grad_close(np32(b.grad), np32(b2.grad), "sharded grad b")


def test_graphed_chamfer_step_and_prefetcher(pp):
    """CUDA-graph step (pipeline.GraphedChamferStep) and the host prefetcher give the same loss and
    gradients as the plain autograd path, also when the two host buffer sets hold different data."""
    from pytorch_points_b200.pipeline import GraphedChamferStep, HostPrefetcher
    pairs = [(uniform_cloud(4, 600, 80 + i).pin_memory(), uniform_cloud(4, 500, 90 + i).pin_memory()) for i in range(2)]
    step = GraphedChamferStep(pairs)
    for it in range(4):
        a0, b0 = pairs[it % 2]
        loss = step.run()
        torch.cuda.synchronize()
        a, b = dev(a0).requires_grad_(True), dev(b0).requires_grad_(True)
        d1, d2, i1, i2 = pp.nndistance(a, b)
        ref = d1.mean() + d2.mean()
        ref.backward()
        assert abs(loss - ref.item()) <= 1e-6 * abs(ref.item())
        assert torch.equal(step.idx1, i1) and torch.equal(step.idx2, i2)
        assert_grad_close(np32(step.grad1), np32(a.grad), "graph grad1")
        assert_grad_close(np32(step.grad2), np32(b.grad), "graph grad2")
    pf = HostPrefetcher("cuda:0")
    pf.prefetch(pairs[0]); pf.prefetch(pairs[1])
    x0, y0 = pf.get(); pf.release()
    x1, y1 = pf.get(); pf.release()
    torch.cuda.synchronize()
    assert torch.equal(x0.cpu(), pairs[0][0]) and torch.equal(y1.cpu(), pairs[1][1])


def test_host_scalar_reader_in_the_autograd_loop(pp):
    """pipeline.HostScalarReader: every step's loss arrives on the host (one step behind, in order) with the
    value loss.item() would have given, while prefetcher, autograd loss and backward keep running; the
    non-contiguous / wrong-dtype input path and the bounds are checked too."""
    from pytorch_points_b200.dist import sharded_chamfer_loss
    from pytorch_points_b200.pipeline import HostPrefetcher, HostScalarReader
    pairs = [(uniform_cloud(4, 700, 180 + i).pin_memory(), uniform_cloud(4, 650, 190 + i).pin_memory()) for i in range(3)]
    pf, rd = HostPrefetcher("cuda", depth=2), HostScalarReader("cuda", depth=2)
    pf.prefetch(pairs[0])
    got, want = [], []
    for it in range(6):
        xd, yd = pf.get()
        x, y = xd.detach().requires_grad_(True), yd.detach().requires_grad_(True)
        loss = sharded_chamfer_loss(x, y, total_batch=4)
        rd.push(loss)
        loss.backward()
        want.append(float(loss.detach().cpu()))
        pf.release()
        pf.prefetch(pairs[(it + 1) % 3])
        if len(rd) > 1:
            got.append(rd.pop())
    while len(rd):
        got.append(rd.pop())
    assert got == want
    # reference value of the first step
    d1, d2, _, _ = pp.nndistance(dev(pairs[0][0]), dev(pairs[0][1]))
    assert abs(got[0] - (d1.mean() + d2.mean()).item()) <= 1e-6 * abs(got[0])
    with pytest.raises(RuntimeError):
        rd.pop()
    rd2 = HostScalarReader("cuda:0", depth=1, numel=2)
    rd2.push(torch.tensor([[1.5], [2.5]], dtype=torch.float64, device="cuda").t())  # reshaped, cast, made contiguous
    with pytest.raises(RuntimeError):
        rd2.push(torch.zeros(2, device="cuda"))
    assert rd2.pop() == [1.5, 2.5]


def test_graphed_chamfer_step_pipelined(pp):
    """submit()/loss(): two steps in flight; each ticket returns the loss of ITS step even though the
    pinned host buffers are rewritten with new clouds as soon as their copy has been consumed."""
    from pytorch_points_b200.pipeline import GraphedChamferStep
    pairs = [(uniform_cloud(4, 700, 180 + i).pin_memory(), uniform_cloud(4, 650, 190 + i).pin_memory()) for i in range(2)]
    clouds = [(uniform_cloud(4, 700, 300 + i), uniform_cloud(4, 650, 400 + i)) for i in range(7)]
    want = []
    for a0, b0 in clouds:
        d1, d2, _, _ = pp.nndistance(dev(a0), dev(b0))
        want.append((d1.mean() + d2.mean()).item())
    step = GraphedChamferStep(pairs)
    for s_ in range(2):
        pairs[s_][0].copy_(clouds[s_][0]); pairs[s_][1].copy_(clouds[s_][1])
    got, tickets = [], []
    for i in range(len(clouds)):
        tickets.append(step.submit())  # enqueues step i and the H2D copy for step i + 1
        if i + 2 < len(clouds):
            # host set i % 2 has been consumed by the copy for step i: refill it for step i + 2
            step.copied[i % 2].synchronize()
            pairs[i % 2][0].copy_(clouds[i + 2][0]); pairs[i % 2][1].copy_(clouds[i + 2][1])
        if len(tickets) > 1:
            got.append(step.loss(tickets.pop(0)))  # loss of step i - 1 while step i runs
    got.append(step.loss(tickets.pop(0)))
    for g, w in zip(got, want):
        assert abs(g - w) <= 1e-6 * abs(w), (got, want)


def test_loss_exchange_single_rank(pp, tmp_path):
    """dist.LossExchange (peer-memory exchange of the loss sums) on a 1-rank group: mailbox creation,
    send/wait kernels, sequence numbering across eager launches and CUDA-graph replays, and the
    graphed Chamfer step built on it.  (W > 1 is exercised by tools/lx_test.py on a multi-GPU box.)"""
    import torch.distributed as dist
    from pytorch_points_b200.dist import LossExchange
    from pytorch_points_b200.pipeline import GraphedChamferStep
    own_group = not dist.is_initialized()
    if own_group:
        dist.init_process_group("nccl", init_method="file://" + str(tmp_path / "pg"), rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
    try:
        lx = LossExchange(torch.device("cuda", 0))
        sums, total = torch.zeros(2, device="cuda"), torch.zeros(2, device="cuda")
        for step in range(5):
            sums.copy_(torch.tensor([1.5 + step, -2.0 * step], device="cuda"))
            lx.send(sums)
            lx.wait(total)
            assert torch.equal(total, sums)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                lx.send(sums)
                lx.wait(total)
        for step in range(4):
            sums.copy_(torch.tensor([9.0 * step, 0.125], device="cuda"))
            torch.cuda.synchronize()
            g.replay()
            torch.cuda.synchronize()
            assert torch.equal(total, sums)
        assert not lx.timed_out()
        pairs = [(uniform_cloud(3, 500, 701).pin_memory(), uniform_cloud(3, 400, 702).pin_memory())]
        step_obj = GraphedChamferStep(pairs, world_size=1, exchange=lx)
        d1, d2, _, _ = pp.nndistance(dev(pairs[0][0]), dev(pairs[0][1]))
        want = (d1.mean() + d2.mean()).item()
        for _ in range(3):
            assert abs(step_obj.run() - want) <= 1e-6 * abs(want)
    finally:
        if own_group:
            dist.destroy_process_group()


def test_labeled_chamfer(pp, oracle_mod):
    a, b = uniform_cloud(2, 700, 27), uniform_cloud(2, 900, 28)
    g = torch.Generator().manual_seed(29)
    la = torch.randint(0, 4, (2, 700, 1), generator=g)
    lb = torch.randint(0, 3, (2, 900, 1), generator=g)  # label 3 has no partner in b
    ga, gb = dev(a).requires_grad_(True), dev(b).requires_grad_(True)
    d1, d2, i1, i2 = pp.labeled_nndistance(ga, gb, dev(la), dev(lb))
    e1, e2, j1, j2 = oracle_mod.chamfer_labeled_fwd(np32(a), np32(b), np32(la.float()), np32(lb.float()))
    assert np.array_equal(np32(i1), j1) and np.array_equal(np32(i2), j2)
    assert np.array_equal(np32(d1), e1) and np.array_equal(np32(d2), e2)
    assert (np32(i1) == -1).any()
    (d1.sum() + d2.sum()).backward()
    o1, o2 = oracle_mod.chamfer_bwd(np32(a), np32(b), np.ones((2, 700), np.float32), np.ones((2, 900), np.float32), j1, j2)
    assert_grad_close(np32(ga.grad), o1, "labeled gradxyz1")
    assert_grad_close(np32(gb.grad), o2, "labeled gradxyz2")


@pytest.mark.parametrize("B,N,M,nl,dup", [(1, 1, 1, 1, False), (2, 33, 5000, 5, True), (3, 4500, 257, 2, False),
                                           (1, 2500, 2500, 40, True), (2, 127, 129, 300, False)])
def test_labeled_chamfer_fast_path_matches_oracle_and_generic(pp, oracle_mod, B, N, M, nl, dup):
    """One-pass kernel with the label mask (pp_chamfer_labeled_fwd with a workspace) against the
    oracle and against the chunk-by-chunk restatement (chamfer_generic=1): ragged sizes, both
    reference-block widths (M <= 4096 / above), duplicated points (ties must resolve to the lowest
    same-label index), labels without a partner on either side (-> idx -1, dist 0)."""
    from pytorch_points_b200 import _C
    a, b = uniform_cloud(B, N, 61), uniform_cloud(B, M, 62)
    if dup:
        a, b = with_duplicates(a), with_duplicates(b)
    g = torch.Generator().manual_seed(63)
    la = torch.randint(0, nl + 1, (B, N, 1), generator=g)  # label nl never appears in b
    lb = torch.randint(-1, nl, (B, M, 1), generator=g)     # label -1 never appears in a
    e1, e2, j1, j2 = oracle_mod.chamfer_labeled_fwd(np32(a), np32(b), np32(la.float()), np32(lb.float()))
    for generic in (0, 1, 0):  # fast, generic, fast again (the key workspace must come back clean)
        _C.set_option("chamfer_generic", generic)
        try:
            d1, d2, i1, i2 = pp.labeled_nndistance(dev(a), dev(b), dev(la), dev(lb))
        finally:
            _C.set_option("chamfer_generic", 0)
        assert np.array_equal(np32(i1), j1) and np.array_equal(np32(i2), j2), "generic=%d" % generic
        assert np.array_equal(np32(d1), e1) and np.array_equal(np32(d2), e2), "generic=%d" % generic
    # the unlabeled path shares the workspace: it must still see it clean
    d1, d2, i1, i2 = pp.nndistance(dev(a), dev(b))
    f1, f2, k1, k2 = oracle_mod.chamfer_fwd(np32(a), np32(b))
    assert np.array_equal(np32(i1), k1) and np.array_equal(np32(i2), k2)
    assert np.array_equal(np32(d1), f1) and np.array_equal(np32(d2), f2)


def test_labeled_chamfer_single_label_equals_unlabeled(pp):
    a, b = dev(with_duplicates(uniform_cloud(2, 3000, 64))), dev(uniform_cloud(2, 5000, 65))
    la, lb = torch.zeros(2, 3000, 1, device="cuda"), torch.zeros(2, 5000, 1, device="cuda")
    got = pp.labeled_nndistance(a, b, la, lb)
    want = pp.nndistance(a, b)
    for x, y in zip(got, want):
        assert torch.equal(x, y)


@pytest.mark.parametrize("B,N,M", [(2, 31, 1001), (3, 1001, 250), (1, 4097, 4099), (2, 2500, 2500)])
def test_chamfer_entry_points_agree(pp, oracle_mod, B, N, M):
    """Plain, labeled and fused-backward entry points on the same clouds: odd cloud strides,
    partial last granules / query groups, duplicated points."""
    from pytorch_points_b200._ext import losses
    a, b = with_duplicates(uniform_cloud(B, N, 81)), with_duplicates(uniform_cloud(B, M, 82))
    g = torch.Generator().manual_seed(83)
    la, lb = torch.randint(0, 3, (B, N, 1), generator=g), torch.randint(1, 4, (B, M, 1), generator=g)
    e = oracle_mod.chamfer_fwd(np32(a), np32(b))
    el = oracle_mod.chamfer_labeled_fwd(np32(a), np32(b), np32(la.float()), np32(lb.float()))
    got = pp.nndistance(dev(a), dev(b))
    gotl = pp.labeled_nndistance(dev(a), dev(b), dev(la), dev(lb))
    d1 = torch.empty(B, N, device="cuda"); d2 = torch.empty(B, M, device="cuda")
    i1 = torch.empty(B, N, dtype=torch.int32, device="cuda"); i2 = torch.empty(B, M, dtype=torch.int32, device="cuda")
    g1, g2 = torch.empty(B, N, 3, device="cuda"), torch.empty(B, M, 3, device="cuda")
    gw = torch.tensor([0.5, 2.0], device="cuda")
    losses.nmdistance_forward_backward_uniform(dev(a), dev(b), d1, d2, i1, i2, None, gw, g1, g2)
    for x, y in zip(got, e):
        assert np.array_equal(np32(x), y)
    for x, y in zip(gotl, el):
        assert np.array_equal(np32(x), y)
    for x, y in zip((d1, d2, i1, i2), e):
        assert np.array_equal(np32(x), y)
    o1, o2 = oracle_mod.chamfer_bwd(np32(a), np32(b), np.full((B, N), 0.5, np.float32), np.full((B, M), 2.0, np.float32), e[2], e[3])
    assert_grad_close(np32(g1), o1, "gradxyz1")
    assert_grad_close(np32(g2), o2, "gradxyz2")


def test_chamfer_unaligned_views(pp, oracle_mod):
    """Clouds that start 4 bytes into an allocation: no kernel may assume 16-byte alignment of
    the base pointers."""
    from pytorch_points_b200._ext import losses
    a, b = uniform_cloud(2, 640, 84), uniform_cloud(2, 512, 85)
    e = oracle_mod.chamfer_fwd(np32(a), np32(b))
    abuf = torch.empty(a.numel() + 1, device="cuda"); bbuf = torch.empty(b.numel() + 1, device="cuda")
    av, bv = abuf[1:].view(2, 640, 3), bbuf[1:].view(2, 512, 3)
    av.copy_(a); bv.copy_(b)
    assert av.data_ptr() % 16 == 4 and av.is_contiguous()
    d1 = torch.empty(2, 640, device="cuda"); d2 = torch.empty(2, 512, device="cuda")
    i1 = torch.empty(2, 640, dtype=torch.int32, device="cuda"); i2 = torch.empty(2, 512, dtype=torch.int32, device="cuda")
    losses.nmdistance_forward(av, bv, d1, d2, i1, i2)
    for x, y in zip((d1, d2, i1, i2), e):
        assert np.array_equal(np32(x), y)


def test_chamfer_fused_sums(pp):
    from pytorch_points_b200._ext import losses
    a, b = dev(uniform_cloud(3, 999, 30)), dev(uniform_cloud(3, 1001, 31))
    d1 = torch.empty(3, 999, device="cuda"); d2 = torch.empty(3, 1001, device="cuda")
    i1 = torch.empty(3, 999, dtype=torch.int32, device="cuda"); i2 = torch.empty(3, 1001, dtype=torch.int32, device="cuda")
    sums = torch.full((2,), 123.0, device="cuda")
    losses.nmdistance_forward(a, b, d1, d2, i1, i2, sums=sums)
    assert torch.allclose(sums[0], d1.sum(), rtol=1e-5) and torch.allclose(sums[1], d2.sum(), rtol=1e-5)


@pytest.mark.parametrize("B,N,M,dup", [(1, 1, 1, False), (3, 999, 1001, True), (2, 5000, 300, False),
                                        (2, 257, 4500, True), (4, 2500, 2500, False)])
def test_chamfer_fused_forward_backward(pp, oracle_mod, B, N, M, dup):
    """pp_chamfer_fwd_bwd_uniform (backward folded into the index-resolving kernel) against the
    oracle: dist/idx bit-exact, gradients within 1e-5; twice in a row on dirty gradient buffers (the
    forward kernel clears them) and followed by a plain forward on the shared key workspace."""
    from pytorch_points_b200._ext import losses
    a, b = uniform_cloud(B, N, 71), uniform_cloud(B, M, 72)
    if dup:
        a, b = with_duplicates(a), with_duplicates(b)
    e1, e2, j1, j2 = oracle_mod.chamfer_fwd(np32(a), np32(b))
    w = (0.25 / (B * N), 3.0 / (B * M))
    o1, o2 = oracle_mod.chamfer_bwd(np32(a), np32(b), np.full((B, N), w[0], np.float32), np.full((B, M), w[1], np.float32), j1, j2)
    ad, bd = dev(a), dev(b)
    d1 = torch.empty(B, N, device="cuda"); d2 = torch.empty(B, M, device="cuda")
    i1 = torch.empty(B, N, dtype=torch.int32, device="cuda"); i2 = torch.empty(B, M, dtype=torch.int32, device="cuda")
    sums = torch.full((2,), 55.0, device="cuda")
    gw = torch.tensor(w, device="cuda")
    g1, g2 = torch.full_like(ad, 7.0), torch.full_like(bd, -7.0)
    for _ in range(2):
        losses.nmdistance_forward_backward_uniform(ad, bd, d1, d2, i1, i2, sums, gw, g1, g2)
        assert np.array_equal(np32(i1), j1) and np.array_equal(np32(i2), j2)
        assert np.array_equal(np32(d1), e1) and np.array_equal(np32(d2), e2)
        assert torch.allclose(sums[0], d1.sum(), rtol=1e-5) and torch.allclose(sums[1], d2.sum(), rtol=1e-5)
        assert_grad_close(np32(g1), o1, "fused gradxyz1")
        assert_grad_close(np32(g2), o2, "fused gradxyz2")
    # same numbers as the separate backward
    h1, h2 = torch.empty_like(ad), torch.empty_like(bd)
    losses.nmdistance_backward_uniform(ad, bd, h1, h2, gw, i1, i2)
    assert_grad_close(np32(g1), np32(h1), "fused vs separate gradxyz1")
    assert_grad_close(np32(g2), np32(h2), "fused vs separate gradxyz2")
    losses.nmdistance_forward(ad, bd, d1, d2, i1, i2)
    assert np.array_equal(np32(i1), j1) and np.array_equal(np32(i2), j2)


def test_chamfer_fused_forward_backward_empty_side(pp):
    from pytorch_points_b200._ext import losses
    a, b = dev(uniform_cloud(2, 10, 73)), torch.empty(2, 0, 3, device="cuda")
    d1 = torch.full((2, 10), 5.0, device="cuda"); d2 = torch.empty(2, 0, device="cuda")
    i1 = torch.full((2, 10), 5, dtype=torch.int32, device="cuda"); i2 = torch.empty(2, 0, dtype=torch.int32, device="cuda")
    g1, g2 = torch.full_like(a, 7.0), torch.empty_like(b)
    losses.nmdistance_forward_backward_uniform(a, b, d1, d2, i1, i2, None, torch.ones(2, device="cuda"), g1, g2)
    assert not d1.any() and not i1.any() and not g1.any()


def test_graphed_step_fused_and_separate_agree(pp):
    from pytorch_points_b200.pipeline import GraphedChamferStep
    a, b = uniform_cloud(4, 700, 74).pin_memory(), uniform_cloud(4, 900, 75).pin_memory()
    fused = GraphedChamferStep([(a, b)], fused_backward=True)
    plain = GraphedChamferStep([(a, b)], fused_backward=False)
    assert fused.launches == 2 and plain.launches == 4
    for _ in range(3):
        lf, lp = fused.run(), plain.run()
        assert abs(lf - lp) <= 1e-6 * abs(lp)
    torch.cuda.synchronize()
    assert_grad_close(np32(fused.grad1), np32(plain.grad1), "graphed fused grad1")
    assert_grad_close(np32(fused.grad2), np32(plain.grad2), "graphed fused grad2")
    assert torch.equal(fused.idx1, plain.idx1) and torch.equal(fused.idx2, plain.idx2)


# --------------------------------------------------------------------------- FPS
FPS_CASES = [
    # (B, N, m, maker, seed_idx)
    (2, 1, 1, uniform_cloud, 0),
    (2, 5, 5, uniform_cloud, 2),
    (2, 100, 30, uniform_cloud, 0),        # bs = 64
    (3, 511, 64, uniform_cloud, 7),        # bs = 256
    (2, 512, 512, uniform_cloud, 0),       # exhaust the cloud
    (2, 1000, 100, sphere_cloud, 3),
    (2, 4096, 256, uniform_cloud, 0),
    (2, 5000, 128, uniform_cloud, 11),     # not a multiple of anything
    (4, 16384, 128, uniform_cloud, 0),     # config 3 cloud size, fewer samples
    (1, 70000, 40, uniform_cloud, 5),      # beyond the register-resident kernel -> streaming path
]


@pytest.mark.parametrize("B,N,m,maker,seed_idx", FPS_CASES)
def test_fps_bit_exact(pp, oracle_mod, B, N, m, maker, seed_idx):
    x = maker(B, N, 3000 + N)
    idx, pts = pp.furthest_point_sample(dev(x), m, NCHW=False, seedIdx=seed_idx)
    want = oracle_mod.fps(np32(x), m, seed=seed_idx)
    assert idx.dtype == torch.int32 and idx.shape == (B, m)
    assert np.array_equal(np32(idx), want)
    assert pts.shape == (B, m, 3)
    assert np.array_equal(np32(pts), np.take_along_axis(np32(x), want[..., None].astype(np.int64), 1))


@pytest.mark.parametrize("maker", [with_duplicates, None])
@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
def test_fps_ties_every_cluster_width(pp, oracle_mod, maker, cluster):
    """Duplicated / lattice clouds make exact ties common; the reference's (k mod 512, k)
    tie key must hold for every cluster width and the npoint > #unique case."""
    from pytorch_points_b200 import _C
    x = with_duplicates(uniform_cloud(2, 4096, 31), 0.3) if maker else lattice_cloud(2, 4096, 32, levels=6)
    want = oracle_mod.fps(np32(x), 300, seed=1)
    _C.set_option("fps_cluster", cluster)
    try:
        idx, _ = pp.furthest_point_sample(dev(x), 300, NCHW=False, seedIdx=1)
    finally:
        _C.set_option("fps_cluster", 0)
    assert np.array_equal(np32(idx), want)


def test_fps_stream_kernel_matches(pp, oracle_mod):
    from pytorch_points_b200 import _C
    x = with_duplicates(uniform_cloud(2, 3000, 33))
    want = oracle_mod.fps(np32(x), 200, seed=0)
    _C.set_option("fps_stream", 1)
    try:
        idx, _ = pp.furthest_point_sample(dev(x), 200, NCHW=False)
    finally:
        _C.set_option("fps_stream", 0)
    assert np.array_equal(np32(idx), want)


def test_fps_temp_contents_and_nchw(pp, oracle_mod):
    from pytorch_points_b200._ext import sampling
    x = uniform_cloud(2, 2000, 34)
    xd = dev(x)
    temp = torch.full((2, 2000), 1e10, device="cuda")
    idx = torch.empty(2, 50, dtype=torch.int32, device="cuda")
    sampling.furthest_sampling(50, 0, xd, temp, idx)
    want_idx, want_temp = oracle_mod.fps(np32(x), 50, return_temp=True)
    assert np.array_equal(np32(idx), want_idx)
    assert np.array_equal(np32(temp), want_temp)   # final running minima, bit for bit
    idx2, pts = pp.furthest_point_sample(xd.transpose(1, 2).contiguous(), 50)  # NCHW=True default
    assert np.array_equal(np32(idx2), want_idx) and pts.shape == (2, 3, 50)


def test_fps_bad_seed_raises(pp):
    with pytest.raises(RuntimeError):
        pp.furthest_point_sample(dev(uniform_cloud(1, 10, 35)), 3, NCHW=False, seedIdx=10)


# --------------------------------------------------------------------------- gather / group
def test_gather_forward_backward(pp, oracle_mod):
    f = uniform_cloud(3, 500, 36, c=7).transpose(1, 2).contiguous()  # (B,C,N)
    g = torch.Generator().manual_seed(37)
    idx = torch.randint(0, 500, (3, 123), generator=g, dtype=torch.int32)
    fd = dev(f).requires_grad_(True)
    out = pp.gather_points(fd, dev(idx))
    assert np.array_equal(np32(out), oracle_mod.gather_fwd(np32(f), np32(idx)))
    w = uniform_cloud(3, 123, 38, c=7).transpose(1, 2).contiguous()
    (out * dev(w)).sum().backward()
    assert_grad_close(np32(fd.grad), oracle_mod.gather_bwd(np32(w), np32(idx), 500), "gather grad")
    # int64 indices are accepted and cast like the reference (operations.py:55)
    out2 = pp.gather_points(dev(f), dev(idx).long())
    assert torch.equal(out2, out.detach())


def test_group_forward_backward(pp, oracle_mod):
    f = uniform_cloud(2, 300, 39, c=5).transpose(1, 2).contiguous()
    g = torch.Generator().manual_seed(40)
    idx = torch.randint(0, 300, (2, 40, 9), generator=g, dtype=torch.int32)
    fd = dev(f).requires_grad_(True)
    out = pp.grouping_operation(fd, dev(idx))
    assert np.array_equal(np32(out), oracle_mod.group_fwd(np32(f), np32(idx)))
    w = torch.rand(2, 5, 40, 9, generator=g)
    (out * dev(w)).sum().backward()
    assert_grad_close(np32(fd.grad), oracle_mod.group_bwd(np32(w), np32(idx), 300), "group grad")


# --------------------------------------------------------------------------- ball_query
BQ_CASES = [
    # (B, N, M, radius, nsample, maker)
    (2, 50, 10, 0.3, 8, uniform_cloud),
    (2, 1000, 200, 0.05, 16, uniform_cloud),     # many empty / short balls
    (2, 4096, 512, 0.2, 32, uniform_cloud),      # > nsample hits, early exit
    (2, 3000, 100, 0.2, 32, sphere_cloud),
    (1, 10, 4, 10.0, 64, uniform_cloud),         # nsample > N
    (2, 777, 33, 0.25, 1, lattice_cloud),        # points exactly on the radius (strict <)
    (2, 2048, 64, 0.0, 4, uniform_cloud),        # r = 0: nothing matches, all zeros
]


@pytest.mark.parametrize("B,N,M,radius,nsample,maker", BQ_CASES)
def test_ball_query_bit_exact(pp, oracle_mod, B, N, M, radius, nsample, maker):
    xyz = maker(B, N, 4000 + N)
    centres = xyz[:, :M].clone() if maker is lattice_cloud else maker(B, M, 5000 + M)
    idx = pp.ball_query(radius, nsample, dev(xyz), dev(centres))
    assert idx.dtype == torch.int32 and idx.shape == (B, M, nsample)
    assert np.array_equal(np32(idx), oracle_mod.ball_query(radius, nsample, np32(xyz), np32(centres)))


def test_query_and_group_module(pp, oracle_mod):
    xyz = uniform_cloud(2, 1024, 41)
    feats = uniform_cloud(2, 1024, 42, c=6).transpose(1, 2).contiguous()
    idx_fps, new_xyz = pp.furthest_point_sample(dev(xyz), 64, NCHW=False)
    grouper = pp.QueryAndGroup(0.2, 16)
    out = grouper(dev(xyz), new_xyz, dev(feats))
    assert out.shape == (2, 9, 64, 16)
    bq = oracle_mod.ball_query(0.2, 16, np32(xyz), np32(new_xyz))
    gx = oracle_mod.group_fwd(np32(xyz.transpose(1, 2).contiguous()), bq) - np32(new_xyz).transpose(0, 2, 1)[..., None]
    gf = oracle_mod.group_fwd(np32(feats), bq)
    assert np.array_equal(np32(out), np.concatenate([gx, gf], 1))


# --------------------------------------------------------------------------- KNN
KNN_CASES = [
    # (B, M, N, k, maker)
    (2, 10, 16, 16, uniform_cloud),       # k == N
    (2, 100, 300, 1, uniform_cloud),
    (2, 2048, 2048, 16, uniform_cloud),   # config 1 shape
    (2, 700, 1500, 16, lattice_cloud),    # ties -> (distance, index) order
    (1, 300, 5000, 32, sphere_cloud),
    (2, 513, 999, 7, uniform_cloud),
    (2, 300, 700, 20, lattice_cloud),     # 16 < k <= 32: one query per thread
    (1, 200, 600, 40, lattice_cloud),     # k > 32: shared-memory list kernel
    (1, 100, 64, 64, uniform_cloud),      # k == N == PP_KNN_MAX_K
    (1, 3000, 20001, 16, uniform_cloud),  # query != points, odd sizes, ordered-sweep path
]


@pytest.mark.parametrize("B,M,N,k,maker", KNN_CASES)
def test_knn_bit_exact(pp, oracle_mod, B, M, N, k, maker):
    p = maker(B, N, 6000 + N)
    q = p[:, :M].clone() if M <= N and maker is not sphere_cloud else maker(B, M, 7000 + M)
    nn, idx, dist = pp.group_knn(k, dev(q), dev(p), NCHW=False)
    ed, ei = oracle_mod.knn(k, np32(q), np32(p))
    assert idx.dtype == torch.int32
    assert np.array_equal(np32(idx), ei)
    assert np.array_equal(np32(dist), ed)
    assert np.array_equal(np32(nn), np.take_along_axis(np32(p)[:, None], ei[..., None].astype(np.int64), 2))


@pytest.mark.parametrize("maker,M,N,k", [(uniform_cloud, 700, 5000, 16), (lattice_cloud, 600, 4500, 16),
                                          (sphere_cloud, 300, 6000, 32), (uniform_cloud, 513, 4097, 5),
                                          (lattice_cloud, 333, 2111, 12),   # k < list width 16, exact ties
                                          (lattice_cloud, 100, 2500, 20),   # full-warp lists (17..32), k < 32
                                          (uniform_cloud, 10, 3000, 1),     # one partial query warp, k = 1
                                          (None, 400, 2600, 16),            # duplicated points: zero distances, index order
                                          (uniform_cloud, 257, 16500, 8)])  # above the one-launch preparation limit
def test_knn_morton_sweep_matches_oracle(pp, oracle_mod, maker, M, N, k):
    """Spatially ordered sweep (forced on): same (distance, original index) order, bit for bit,
    including exact ties (lattice) and query != points."""
    from pytorch_points_b200 import _C
    if maker is None:
        p = with_duplicates(uniform_cloud(2, N, 70), 0.3)
        q = p[:, :M].clone()
    else:
        p = maker(2, N, 70)
        q = maker(2, M, 71)
    ed, ei = oracle_mod.knn(k, np32(q), np32(p))
    _C.set_option("knn_morton", 1)
    try:
        _, idx, dist = pp.group_knn(k, dev(q), dev(p), NCHW=False)
        _, idx_s, dist_s = pp.group_knn(k, dev(p), dev(p), NCHW=False)   # self-KNN shares the sort
    finally:
        _C.set_option("knn_morton", -1)
    assert np.array_equal(np32(idx), ei) and np.array_equal(np32(dist), ed)
    ns = min(800, N)
    ed2, ei2 = oracle_mod.knn(k, np32(p[:1, :ns]), np32(p[:1]))
    assert np.array_equal(np32(idx_s[:1, :ns]), ei2) and np.array_equal(np32(dist_s[:1, :ns]), ed2)


KNN_TC_CASES = [
    # (name, B, M, N, k): the tensor-core path (knn_tc.cu) forced on
    ("uniform", 2, 2048, 2048, 16), ("ragged", 3, 1000, 2500, 16), ("sphere", 2, 3000, 3000, 8),
    ("duplicates", 2, 2048, 2048, 16), ("lattice", 2, 1500, 1500, 16), ("offset", 2, 2048, 2048, 16),
    ("tiny", 2, 2048, 2048, 16), ("fewpoints", 2, 300, 40, 32), ("k1", 2, 2048, 2048, 1), ("k32", 1, 4096, 4096, 32),
    ("allequal", 1, 600, 600, 16), ("clusters", 2, 4096, 4096, 16), ("above16k", 1, 300, 16500, 8),
    ("onequery", 1, 1, 2048, 3), ("oddtile", 1, 129, 4000, 17), ("kequalsn", 2, 50, 24, 24), ("manyclouds", 40, 130, 260, 5),
]


def _knn_tc_input(name, B, n, seed):
    if name == "sphere":
        return sphere_cloud(B, n, seed)
    if name == "duplicates":
        return with_duplicates(uniform_cloud(B, n, seed))
    if name == "lattice":
        return lattice_cloud(B, n, seed)
    if name == "offset":
        return uniform_cloud(B, n, seed) + 1000.0
    if name == "tiny":
        return uniform_cloud(B, n, seed) * 1e-4
    if name == "allequal":
        return torch.zeros(B, n, 3) + 0.25
    if name == "clusters":  # two clusters of 1e-3 a hundred units apart: everything is within the approximation's error
        g = torch.Generator().manual_seed(seed)
        return uniform_cloud(B, n, seed) * 1e-3 + 100.0 * torch.randint(0, 2, (B, n, 1), generator=g).float()
    return uniform_cloud(B, n, seed)


@pytest.mark.parametrize("name,B,M,N,k", KNN_TC_CASES)
def test_knn_tensor_core_path_matches_oracle(pp, oracle_mod, name, B, M, N, k):
    """knn_tc.cu (tcgen05 flagging pass + exact resolution), forced on: distances and indices equal the
    oracle's bit for bit -- ties, duplicates, lattices, far-from-origin and degenerate inputs, query != points."""
    from pytorch_points_b200 import _C
    p = _knn_tc_input(name, B, N, 11)
    q = p if M == N else _knn_tc_input(name, B, M, 12)
    ed, ei = oracle_mod.knn(k, np32(q), np32(p))
    pd = dev(p)
    qd = pd if q is p else dev(q)
    _C.set_option("knn_tc", 1)
    try:
        _, idx, dist = pp.group_knn(k, qd, pd, NCHW=False)
    finally:
        _C.set_option("knn_tc", -1)
    assert np.array_equal(np32(idx), ei) and np.array_equal(np32(dist), ed)


def test_knn_points_adaptor_at_the_callers_list_width(pp, oracle_mod):
    """`ops.knn_points(x, x, K=k+1, return_nn=True)` with k = 16 (network/layers.py:52) and K = nn_size = 20
    (geo_operations.py:112): lists wider than 16 take the tensor-core path automatically -- same bits as the
    oracle, self dropped by the caller's slice."""
    x = uniform_cloud(2, 4096, 321)
    for K in (17, 20):
        ed, ei = oracle_mod.knn(K, np32(x), np32(x))
        d, i, nn = pp.knn_points(dev(x), dev(x), K=K, return_nn=True)
        assert i.dtype == torch.int64
        assert np.array_equal(np32(i), ei.astype(np.int64)) and np.array_equal(np32(d), ed)
        assert np.array_equal(np32(i[:, :, 0]), np.broadcast_to(np.arange(4096), (2, 4096)))  # self first
        assert np.array_equal(np32(nn), np.take_along_axis(np32(x)[:, None], ei[..., None].astype(np.int64), 2))
    q = uniform_cloud(2, 1500, 322)
    ed, ei = oracle_mod.knn(17, np32(q), np32(x))
    d, i = pp.knn_points(dev(q), dev(x), K=17)[:2]
    assert np.array_equal(np32(i), ei.astype(np.int64)) and np.array_equal(np32(d), ed)


def test_knn_tensor_core_path_nonfinite_points(pp):
    """Non-finite points are never neighbours and a non-finite query gets (inf, -1): same as the ordered sweep."""
    from pytorch_points_b200 import _C
    from pytorch_points_b200._ext import sampling
    p = uniform_cloud(1, 2048, 5)
    p[0, 7] = float("nan")
    p[0, 100, 1] = float("inf")
    pd = dev(p)
    out = []
    for tc in (1, 0):
        _C.set_option("knn_tc", tc)
        try:
            out.append(sampling.knn(8, pd, pd))
        finally:
            _C.set_option("knn_tc", -1)
    assert torch.equal(out[0][1], out[1][1])
    assert torch.equal(torch.nan_to_num(out[0][0], 7.0), torch.nan_to_num(out[1][0], 7.0))
    assert out[0][1][0, 7].tolist() == [-1] * 8


def test_knn_shared_list_kernel_matches(pp, oracle_mod):
    from pytorch_points_b200 import _C
    p = with_duplicates(uniform_cloud(2, 3000, 60), 0.2)
    ed, ei = oracle_mod.knn(16, np32(p[:, :500]), np32(p))
    _C.set_option("knn_smem_lists", 1)
    try:
        _, idx, dist = pp.group_knn(16, dev(p[:, :500].contiguous()), dev(p), NCHW=False)
    finally:
        _C.set_option("knn_smem_lists", 0)
    assert np.array_equal(np32(idx), ei) and np.array_equal(np32(dist), ed)
    _, idx, dist = pp.group_knn(16, dev(p[:, :500].contiguous()), dev(p), NCHW=False)
    assert np.array_equal(np32(idx), ei) and np.array_equal(np32(dist), ed)


def test_knn_nchw_and_knn_points_adaptor(pp, oracle_mod):
    p = uniform_cloud(2, 600, 43)
    ed, ei = oracle_mod.knn(5, np32(p), np32(p))
    nn, idx, dist = pp.group_knn(5, dev(p).transpose(1, 2).contiguous(), dev(p).transpose(1, 2).contiguous())
    assert nn.shape == (2, 3, 600, 5) and np.array_equal(np32(idx), ei)
    d, i, n2 = pp.knn_points(dev(p), dev(p), K=5, return_nn=True)
    assert i.dtype == torch.int64 and np.array_equal(np32(i), ei) and np.array_equal(np32(d), ed)
    assert (np32(i)[..., 0] == np.arange(600)[None]).all()  # self is the nearest neighbour
    assert n2.shape == (2, 600, 5, 3)


def test_knn_generic_dim_and_backward(pp, oracle_mod):
    p, q = uniform_cloud(2, 200, 44, c=4), uniform_cloud(2, 50, 45, c=4)
    pd, qd = dev(p).requires_grad_(True), dev(q).requires_grad_(True)
    nn, idx, dist = pp.group_knn(6, qd, pd, NCHW=False)
    ed, ei = oracle_mod.knn(6, np32(q), np32(p))
    assert np.array_equal(np32(idx), ei) and np.array_equal(np32(dist), ed)
    dist.sum().backward()
    p2, q2 = dev(p).requires_grad_(True), dev(q).requires_grad_(True)
    D = ((q2[:, :, None] - p2[:, None]) ** 2).sum(-1)
    D.topk(6, dim=2, largest=False)[0].sum().backward()
    assert_grad_close(np32(qd.grad), np32(q2.grad), "knn grad query")
    assert_grad_close(np32(pd.grad), np32(p2.grad), "knn grad points")


def test_knn_k_too_large_raises(pp):
    with pytest.raises(RuntimeError):
        pp.group_knn(20, dev(uniform_cloud(1, 5, 46)), dev(uniform_cloud(1, 10, 47)), NCHW=False)


def test_three_nn(pp, oracle_mod):
    from pytorch_points_b200._ext import sampling
    u, k = uniform_cloud(2, 777, 48), uniform_cloud(2, 1300, 49)
    d, i = sampling.three_nn(dev(u), dev(k))
    ed, ei = oracle_mod.three_nn(np32(u), np32(k))
    assert np.array_equal(np32(i), ei) and np.array_equal(np32(d), ed)


def test_three_interpolate(pp, oracle_mod):
    u, k = uniform_cloud(2, 500, 50), uniform_cloud(2, 300, 51)
    dist, idx = pp.three_nn(dev(u), dev(k))
    w = 1.0 / (dist + 1e-8)
    w = (w / w.sum(dim=2, keepdim=True)).contiguous()
    f = uniform_cloud(2, 300, 52, c=5).transpose(1, 2).contiguous()
    fd = dev(f).requires_grad_(True)
    out = pp.three_interpolate(fd, idx, w)
    assert np.array_equal(np32(out), oracle_mod.three_interpolate_fwd(np32(f), np32(idx), np32(w)))
    go = uniform_cloud(2, 500, 53, c=5).transpose(1, 2).contiguous()
    out.backward(dev(go))
    assert_grad_close(np32(fd.grad), oracle_mod.three_interpolate_bwd(np32(go), np32(idx), np32(w), 300), "interp grad")


def test_pointnet2_sa_and_fp_stage_end_to_end(pp):
    """the PointNet++ set-abstraction stage of the reference
    (network/pointnet2_modules.py:21-54: FPS -> gather -> ball_query -> group -> shared MLP -> max
    pool) followed by a feature-propagation stage (:115-153: three_nn -> three_interpolate -> MLP),
    written against OUR ops exactly the way the reference's modules call theirs; forward and
    backward run and the gradient reaches the input features and the MLP weights."""
    torch.manual_seed(0)
    B, N, C, npoint = 2, 2048, 8, 256
    xyz = dev(uniform_cloud(B, N, 54))
    feats = dev(uniform_cloud(B, N, 55, c=C)).transpose(1, 2).contiguous().requires_grad_(True)
    mlp = torch.nn.Sequential(torch.nn.Conv2d(C + 3, 16, 1), torch.nn.ReLU(), torch.nn.Conv2d(16, 32, 1)).cuda()
    fp_mlp = torch.nn.Sequential(torch.nn.Conv1d(32 + C, 16, 1), torch.nn.ReLU()).cuda()
    # --- SA stage
    xyz_flipped = xyz.transpose(1, 2).contiguous()
    idx = pp.furthest_point_sample(xyz, npoint, NCHW=False)[0]
    new_xyz = pp.gather_points(xyz_flipped, idx).transpose(1, 2).contiguous()
    grouped = pp.QueryAndGroup(0.2, 16)(xyz, new_xyz, feats)            # (B, C+3, npoint, nsample)
    new_feats = torch.nn.functional.max_pool2d(mlp(grouped), kernel_size=[1, grouped.size(3)]).squeeze(-1)
    assert new_xyz.shape == (B, npoint, 3) and new_feats.shape == (B, 32, npoint)
    # --- FP stage: propagate the 256 coarse features back to the 2048 points
    dist, nn_idx = pp.three_nn(xyz, new_xyz)
    w = 1.0 / (dist + 1e-8)
    w = w / torch.sum(w, dim=2, keepdim=True)
    interp = pp.three_interpolate(new_feats.contiguous(), nn_idx, w.contiguous())
    out = fp_mlp(torch.cat([interp, feats], dim=1))
    assert out.shape == (B, 16, N)
    out.square().mean().backward()
    assert feats.grad is not None and torch.isfinite(feats.grad).all() and feats.grad.abs().sum() > 0
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in mlp.parameters())


# --------------------------------------------------------------------------- fused SA stages
@pytest.mark.parametrize("B,N,m,seed_idx,stream", [(2, 100, 30, 0, 0), (3, 2048, 256, 5, 0), (2, 16384, 300, 0, 0),
                                                   (2, 9000, 100, 17, 0), (2, 3000, 64, 1, 1)])
def test_fps_gather_fused(pp, oracle_mod, B, N, m, seed_idx, stream):
    """pp_fps_gather: the sampling kernel also emits the sampled coordinates (what
    furthest_point_sample's gather_points + transposes produce, geo_operations.py:60-63)."""
    from pytorch_points_b200 import _C
    xyz = uniform_cloud(B, N, 600 + N)
    _C.set_option("fps_stream", stream)
    try:
        idx, new_xyz = pp.furthest_point_sample(dev(xyz), m, NCHW=False, seedIdx=seed_idx)
        idx_t, new_t = pp.furthest_point_sample(dev(xyz).transpose(1, 2).contiguous(), m, NCHW=True, seedIdx=seed_idx)
    finally:
        _C.set_option("fps_stream", 0)
    want = oracle_mod.fps(np32(xyz), m, seed=seed_idx)
    assert np.array_equal(np32(idx), want) and np.array_equal(np32(idx_t), want)
    gathered = np.take_along_axis(np32(xyz), want[..., None].astype(np.int64), axis=1)
    assert new_xyz.shape == (B, m, 3) and np.array_equal(np32(new_xyz), gathered)
    assert new_t.shape == (B, 3, m) and np.array_equal(np32(new_t), gathered.transpose(0, 2, 1))


def test_fps_gather_backward_scatters_into_xyz(pp):
    xyz = dev(uniform_cloud(2, 500, 611)).requires_grad_(True)
    idx, new_xyz = pp.furthest_point_sample(xyz, 40, NCHW=False)
    w = torch.rand_like(new_xyz)
    (new_xyz * w).sum().backward()
    want = torch.zeros_like(xyz)
    want.scatter_add_(1, idx.long()[..., None].expand(-1, -1, 3), w)
    assert torch.equal(xyz.grad, want)  # FPS indices are distinct here: no summation-order freedom


QG_CASES = [
    # (B, N, M, C, radius, nsample, use_xyz, maker)
    (2, 1024, 64, 6, 0.2, 16, True, uniform_cloud),
    (2, 4096, 203, 13, 0.2, 32, True, uniform_cloud),     # M not a multiple of the CTA's 8 centres
    (2, 3000, 100, 0, 0.2, 32, True, sphere_cloud),       # no features
    (2, 1000, 77, 5, 0.05, 16, False, uniform_cloud),     # features only; many empty / padded balls
    (1, 2000, 50, 3, 0.3, 64, True, uniform_cloud),       # nsample > one warp
    (2, 777, 33, 4, 0.25, 5, True, lattice_cloud),        # exact-radius ties, odd nsample
    (16, 16384, 1024, 16, 0.2, 32, True, uniform_cloud),  # config 3 shape
]


def _oracle_query_group(oracle_mod, xyz, centres, feats, radius, nsample, use_xyz):
    bq = oracle_mod.ball_query(radius, nsample, np32(xyz), np32(centres))
    parts = []
    if use_xyz:
        parts.append(oracle_mod.group_fwd(np32(xyz.transpose(1, 2).contiguous()), bq)
                     - np32(centres).transpose(0, 2, 1)[..., None])
    if feats is not None:
        parts.append(oracle_mod.group_fwd(np32(feats), bq))
    return np.concatenate(parts, 1), bq


@pytest.mark.parametrize("B,N,M,C,radius,nsample,use_xyz,maker", QG_CASES)
def test_query_and_group_fused_bit_exact(pp, oracle_mod, B, N, M, C, radius, nsample, use_xyz, maker):
    xyz = maker(B, N, 620 + N)
    centres = xyz[:, :M].clone() if maker is lattice_cloud else maker(B, M, 630 + M)
    feats = uniform_cloud(B, N, 640 + N, c=C).transpose(1, 2).contiguous() if C else None
    out, idx = pp.query_and_group(dev(xyz), dev(centres), dev(feats) if C else None, radius, nsample, use_xyz)
    assert out.shape == (B, (3 if use_xyz else 0) + C, M, nsample) and idx.dtype == torch.int32
    # op-by-op path through the single kernels must agree exactly at every size
    ref_module = pp.QueryAndGroup(radius, nsample, use_xyz=use_xyz, fused=False)
    assert torch.equal(out, ref_module(dev(xyz), dev(centres), dev(feats) if C else None))
    assert torch.equal(idx, pp.ball_query(radius, nsample, dev(xyz), dev(centres)))
    if C:
        # features staged point-major (what a set-abstraction level does once for all its scales): same bits
        staged = pp.stage_features(dev(feats))
        assert staged.shape == (B, N, C) and torch.equal(staged, dev(feats).transpose(1, 2))
        out_pm, idx_pm = pp.query_and_group(dev(xyz), dev(centres), dev(feats), radius, nsample, use_xyz, staged)
        assert torch.equal(out_pm, out) and torch.equal(idx_pm, idx)
    if B * N * M <= 2 * 4096 * 256:
        want, bq = _oracle_query_group(oracle_mod, xyz, centres, feats, radius, nsample, use_xyz)
        assert np.array_equal(np32(idx), bq)
        assert np.array_equal(np32(out), want)


@pytest.mark.parametrize("C,use_xyz", [(6, True), (0, True), (5, False)])
def test_query_and_group_fused_backward(pp, oracle_mod, C, use_xyz):
    B, N, M, ns = 2, 1500, 120, 16
    xyz = uniform_cloud(B, N, 651)
    centres = uniform_cloud(B, M, 652)
    feats = uniform_cloud(B, N, 653, c=C).transpose(1, 2).contiguous() if C else None
    xd = dev(xyz).requires_grad_(True)
    cd = dev(centres).requires_grad_(True)
    fd = dev(feats).requires_grad_(True) if C else None
    out, idx = pp.query_and_group(xd, cd, fd, 0.2, ns, use_xyz, pp.stage_features(fd) if C else None)
    go = torch.rand_like(out)
    out.backward(go)
    gon, bq = np32(go), np32(idx)
    x0 = 3 if use_xyz else 0
    if C:
        assert_grad_close(np32(fd.grad), oracle_mod.group_bwd(np.ascontiguousarray(gon[:, x0:]), bq, N), "grad features")
    if use_xyz:
        gx = oracle_mod.group_bwd(np.ascontiguousarray(gon[:, :3]), bq, N).transpose(0, 2, 1)
        assert_grad_close(np32(xd.grad), gx, "grad xyz")
        assert_grad_close(np32(cd.grad), -gon[:, :3].astype(np.float64).sum(-1).transpose(0, 2, 1), "grad new_xyz")
    else:
        assert xd.grad is None and cd.grad is None
    # and the op-by-op autograd path gives the same gradients
    x2 = dev(xyz).requires_grad_(True)
    c2 = dev(centres).requires_grad_(True)
    f2 = dev(feats).requires_grad_(True) if C else None
    pp.QueryAndGroup(0.2, ns, use_xyz=use_xyz, fused=False)(x2, c2, f2).backward(go)
    if C:
        assert_grad_close(np32(fd.grad), np32(f2.grad), "fused vs composed: features")
    if use_xyz:
        assert_grad_close(np32(xd.grad), np32(x2.grad), "fused vs composed: xyz")
        assert_grad_close(np32(cd.grad), np32(c2.grad), "fused vs composed: new_xyz")


def test_pointnet_sa_module_matches_op_by_op_restatement(pp, oracle_mod):
    """PointnetSAModuleMSG (network/pointnet2_modules.py:21-93) on the fused kernels vs the
    reference's op-by-op sequence on the single kernels with the same weights; centres are also
    checked against the oracle's FPS."""
    torch.manual_seed(1)
    B, N, C, npoint = 2, 4096, 8, 256
    xyz_c = uniform_cloud(B, N, 661)
    xyz = dev(xyz_c)
    feats = dev(uniform_cloud(B, N, 662, c=C)).transpose(1, 2).contiguous().requires_grad_(True)
    sa = pp.PointnetSAModuleMSG(npoint=npoint, radii=[0.1, 0.2], nsamples=[16, 32], mlps=[[C, 16, 32], [C, 16, 48]]).cuda()
    new_xyz, new_feats = sa(xyz, feats)
    assert new_xyz.shape == (B, npoint, 3) and new_feats.shape == (B, 80, npoint)
    want_idx = oracle_mod.fps(np32(xyz_c), npoint)
    assert np.array_equal(np32(new_xyz), np.take_along_axis(np32(xyz_c), want_idx[..., None].astype(np.int64), axis=1))
    new_feats.square().mean().backward()
    g_fused = feats.grad.clone()
    w_fused = [p.grad.clone() for p in sa.parameters()]
    # --- restatement: FPS, gather_points, ball_query, grouping_operation x2, cat (reference sequence)
    feats.grad = None
    sa.zero_grad()
    idx = pp.FurthestPointSampling.apply(xyz, npoint, 0)
    ctr = pp.gather_points(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
    assert torch.equal(ctr, new_xyz)
    outs = []
    for g, mlp in zip(sa.groupers, sa.mlps):
        grouped = pp.QueryAndGroup(g.radius, g.nsample, use_xyz=True, fused=False)(xyz, ctr, feats)
        y = mlp(grouped)
        outs.append(torch.nn.functional.max_pool2d(y, kernel_size=[1, y.size(3)]).squeeze(-1))
    ref_feats = torch.cat(outs, dim=1)
    assert torch.allclose(new_feats, ref_feats, rtol=1e-6, atol=1e-7)
    ref_feats.square().mean().backward()
    assert_grad_close(np32(g_fused), np32(feats.grad), "SA module: grad features")
    for a, b_ in zip(w_fused, [p.grad for p in sa.parameters()]):
        assert torch.allclose(a, b_, rtol=1e-4, atol=1e-6)


def test_pointnet_sa_group_all_and_fp_module(pp, oracle_mod):
    """npoint=None -> GroupAll (pointnet2_utils.py:127-150); PointnetFPModule
    (pointnet2_modules.py:115-153) against an oracle three_nn / three_interpolate restatement."""
    torch.manual_seed(2)
    B, N, C, m = 2, 1024, 6, 128
    xyz_c, known_c = uniform_cloud(B, N, 671), uniform_cloud(B, m, 672)
    xyz, known = dev(xyz_c), dev(known_c)
    feats = dev(uniform_cloud(B, N, 673, c=C)).transpose(1, 2).contiguous()
    sa = pp.PointnetSAModule(mlp=[C, 32], npoint=None).cuda()
    new_xyz, glob = sa(xyz, feats)
    assert new_xyz is None and glob.shape == (B, 32, 1)
    known_feats = dev(uniform_cloud(B, m, 674, c=10)).transpose(1, 2).contiguous().requires_grad_(True)
    fp = pp.PointnetFPModule(mlp=[10 + C, 24], normalization=None).cuda()
    out = fp(xyz, known, feats, known_feats)
    assert out.shape == (B, 24, N)
    d2, i3 = oracle_mod.three_nn(np32(xyz_c), np32(known_c))
    dist = torch.sqrt(torch.from_numpy(d2))
    w = 1.0 / (dist + 1e-8)
    w = w / torch.sum(w, dim=2, keepdim=True)
    interp = torch.from_numpy(oracle_mod.three_interpolate_fwd(np32(known_feats), i3, w.numpy()))
    # the interpolation itself (weights formed on the device vs on the host: few-ulp differences)
    dist_d, idx_d = pp.three_nn(xyz, known)
    assert np.array_equal(np32(idx_d), i3)
    wd = 1.0 / (dist_d + 1e-8)
    wd = wd / torch.sum(wd, dim=2, keepdim=True)
    assert torch.allclose(pp.three_interpolate(known_feats.detach(), idx_d, wd.contiguous()), dev(interp), rtol=1e-5, atol=1e-6)
    # through the MLP (cuDNN may use TF32 for the 1x1 convolution: loose tolerance)
    want = fp.mlp(torch.cat([dev(interp), feats], dim=1).unsqueeze(-1)).squeeze(-1)
    assert torch.allclose(out, want, rtol=5e-3, atol=5e-3)
    out.sum().backward()
    assert known_feats.grad is not None and torch.isfinite(known_feats.grad).all()


# --------------------------------------------------------------------------- KNN callers (N4)
def _brute_knn(points, k):
    d = torch.cdist(points.double(), points.double())
    return d.topk(k + 1, dim=-1, largest=False).indices[:, :, 1:]


@pytest.mark.parametrize("nsample", [64, 1])
def test_sampled_dense_edge_conv_on_device(pp, oracle_mod, nsample):
    """SampledDenseEdgeConv (network/layers.py:85-133) on the real operators: centres = FPS order of
    the oracle (or the point nearest the centroid), their features gathered, neighbours = brute-force
    k nearest in feature space with the centre itself dropped."""
    torch.manual_seed(4)
    B, C, N, k = 2, 3, 1500, 6
    xyz_c = uniform_cloud(B, N, 811)
    xyz = dev(xyz_c).transpose(1, 2).contiguous()
    x = xyz.clone().requires_grad_(True)  # features = coordinates: distinct points, no feature ties
    conv = pp.SampledDenseEdgeConv(C, 10, n=3, k=k).cuda()
    y, sxyz, sidx = conv(x, nsample, xyz)
    assert y.shape == (B, conv.out_channels, nsample) and sxyz.shape == (B, 3, nsample) and sidx.shape == (B, nsample)
    if nsample == 1:
        want = ((xyz - xyz.mean(-1, keepdim=True)) ** 2).sum(1).argmin(-1, keepdim=True)
    else:
        want = dev(torch.from_numpy(oracle_mod.fps(np32(xyz_c), nsample).astype(np.int64)))
    assert torch.equal(sidx.long(), want)
    assert torch.equal(sxyz, torch.gather(xyz, 2, want.unsqueeze(1).expand(B, 3, nsample)))
    centres = torch.gather(x.detach(), 2, want.unsqueeze(1).expand(B, C, nsample))
    edge, eidx = conv.get_local_graph(centres, x.detach(), k)
    d = torch.cdist(centres.transpose(1, 2).double(), x.detach().transpose(1, 2).double())
    assert torch.equal(eidx, d.topk(k + 1, dim=-1, largest=False).indices[:, :, 1:])
    assert edge.shape == (B, 2 * C, nsample, k)
    y.square().mean().backward()
    assert torch.isfinite(x.grad).all() and x.grad.abs().sum() > 0


def test_knn_callers_laplacian_edgeconv_and_losses(pp):
    """the snapshot's pytorch3d.knn_points callers on this repo's KNN --
    DenseEdgeConv.get_local_graph (layers.py:41-62), pointUniformLaplacian / batch_normals
    (geo_operations.py:88-152) and the point regularisers (model_loss.py:73-163,326-398) --
    against brute-force torch restatements."""
    torch.manual_seed(3)
    B, N, k = 2, 3000, 8
    pts = dev(uniform_cloud(B, N, 801))
    want_idx = _brute_knn(pts, k)
    # Laplacian
    lap, idx = pp.pointUniformLaplacian(pts, nn_size=k)
    assert torch.equal(idx, want_idx)
    nb = torch.gather(pts.unsqueeze(1).expand(B, N, N, 3), 2, want_idx.unsqueeze(-1).expand(B, N, k, 3))
    assert torch.allclose(lap, pts - nb.mean(2), rtol=1e-5, atol=1e-6)
    lap2, _ = pp.pointUniformLaplacian(pts, knn_idx=idx)
    assert torch.allclose(lap, lap2, rtol=1e-6, atol=1e-7)
    # edge convolution graph + module forward/backward
    x = pts.transpose(1, 2).contiguous().requires_grad_(True)
    conv = pp.DenseEdgeConv(3, 12, n=3, k=k).cuda()
    edge, eidx = conv.get_local_graph(x, k)
    assert edge.shape == (B, 6, N, k) and torch.equal(eidx, want_idx)
    assert torch.allclose(edge[:, 3:], nb.permute(0, 3, 1, 2) - x.unsqueeze(-1), rtol=1e-5, atol=1e-6)
    y, _ = conv(x)
    assert y.shape == (B, conv.out_channels, N)
    y.square().mean().backward()
    assert x.grad is not None and torch.isfinite(x.grad).all() and x.grad.abs().sum() > 0
    # regularisers
    l1 = torch.nn.L1Loss()
    assert pp.PointEdgeLengthLoss(k, l1)(pts, pts).item() == 0.0
    assert abs(pp.PointStretchLoss(k)(pts, 2 * pts).item() - 1.0) < 1e-5
    assert pp.PointStretchLoss(k)(pts, 0.5 * pts).item() == 0.0
    assert pp.PointLaplacianLoss(k, l1)(pts, pts).item() == 0.0
    d2 = ((nb - pts.unsqueeze(2)) ** 2).sum(-1)
    rep = torch.where(d2 < 0.03 ** 2, 1 / torch.sqrt(d2 + 1e-4), torch.zeros_like(d2)).mean()
    assert abs(pp.SimplePointRepulsionLoss(k, 0.03)(pts).item() - rep.item()) <= 1e-5 * max(rep.item(), 1e-6)
    moved = pts.clone().requires_grad_(True)
    pp.SimplePointRepulsionLoss(k, 0.05)(moved).backward()
    assert torch.isfinite(moved.grad).all()
    # PCA normals: a noisy plane z ~ 0 has normals +-z; identical clouds give zero normal loss
    plane = pts.clone()
    plane[..., 2] *= 1e-3
    normals, nidx = pp.batch_normals(plane, nn_size=12, NCHW=False)
    assert normals.shape == (B, N, 3) and nidx.shape == (B, N, 12)
    assert (normals[..., 2].abs() > 0.99).float().mean().item() > 0.99
    assert pp.NormalLoss(nn_size=12, reduction="none")(plane, plane).abs().max().item() < 1e-4


# --------------------------------------------------------------------------- full-size properties
def test_chamfer_target_shape_properties(pp, oracle_mod):
    """B=32, N=M=8192 (north-star target) is too big for the CPU oracle inside a unit test:
    check a sampled subset of rows against the oracle plus size-independent invariants."""
    B, N = 32, 8192
    a, b = uniform_cloud(B, N, 50), uniform_cloud(B, N, 51)
    ad, bd = dev(a), dev(b)
    d1, d2, i1, i2 = pp.nndistance(ad, bd)
    # invariant 1: dist equals the recomputed distance to the reported neighbour, bit for bit
    nb = torch.gather(bd, 1, i1.long().unsqueeze(-1).expand(B, N, 3))
    t = nb - ad
    rec = torch.addcmul(torch.addcmul(t[..., 0] * t[..., 0], t[..., 1], t[..., 1]), t[..., 2], t[..., 2])
    assert torch.allclose(rec, d1, rtol=1e-6, atol=0)
    # invariant 2: symmetry -- swapping the clouds swaps the outputs
    s1, s2, j1, j2 = pp.nndistance(bd, ad)
    assert torch.equal(s1, d2) and torch.equal(s2, d1) and torch.equal(j1, i2) and torch.equal(j2, i1)
    # oracle on 2 of the 32 clouds
    for bb in (0, 31):
        e1, e2, k1, k2 = oracle_mod.chamfer_fwd(np32(a[bb:bb + 1]), np32(b[bb:bb + 1]))
        assert np.array_equal(np32(i1[bb:bb + 1]), k1) and np.array_equal(np32(i2[bb:bb + 1]), k2)
        assert np.array_equal(np32(d1[bb:bb + 1]), e1) and np.array_equal(np32(d2[bb:bb + 1]), e2)


def test_fps_config3_shape(pp, oracle_mod):
    """Config 3: B=16, 16384 -> 1024; the oracle checks 2 clouds, the rest by invariants."""
    x = uniform_cloud(16, 16384, 52)
    idx, pts = pp.furthest_point_sample(dev(x), 1024, NCHW=False)
    want = oracle_mod.fps(np32(x[:2]), 1024)
    assert np.array_equal(np32(idx[:2]), want)
    srt = torch.sort(idx.long(), dim=1)[0]
    assert (srt[:, 1:] != srt[:, :-1]).all()           # no repeats on a cloud of distinct points
    bq = pp.ball_query(0.2, 32, dev(x), pts)
    assert np.array_equal(np32(bq[:1]), oracle_mod.ball_query(0.2, 32, np32(x[:1]), np32(pts[:1])))
