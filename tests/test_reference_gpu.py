"""GPU three-way check: the REFERENCE's own kernels (unmodified sources compiled for sm_100a by
oracle/build_ref.sh into oracle/_ref/, which travels to the GPU box) vs the CPU oracle vs our
kernels, on the same inputs.  This is what pins the oracle (the reference ships no golden
vectors).  Skipped when oracle/_ref/*.so is absent."""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import lattice_cloud, np32, sphere_cloud, uniform_cloud, with_duplicates

pytestmark = pytest.mark.gpu

REF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


@pytest.fixture(scope="module")
def ref():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(os.path.join(REF_DIR, "ref_losses.so")):
        pytest.skip("oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)")
    sys.path.insert(0, REF_DIR)
    import ref_losses
    import ref_sampling
    return ref_losses, ref_sampling


@pytest.fixture(scope="module")
def pp():
    from pytorch_points_b200 import network
    return network


def ref_chamfer(rl, a, b):
    B, N, _ = a.shape
    M = b.shape[1]
    d1 = torch.zeros(B, N, device="cuda"); d2 = torch.zeros(B, M, device="cuda")
    i1 = torch.zeros(B, N, dtype=torch.int32, device="cuda"); i2 = torch.zeros(B, M, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    rl.nmdistance_forward(a, b, d1, d2, i1, i2)  # legacy default stream
    torch.cuda.synchronize()
    return d1, d2, i1, i2


@pytest.mark.parametrize("B,N,M,maker", [(2, 300, 257, uniform_cloud), (4, 2500, 2500, uniform_cloud),
                                          (2, 3000, 1000, sphere_cloud), (2, 2048, 2048, lattice_cloud),
                                          (2, 8192, 8192, uniform_cloud)])
def test_chamfer_three_way(ref, pp, oracle_mod, B, N, M, maker):
    rl, _ = ref
    a, b = maker(B, N, 201), maker(B, M, 202)
    ad, bd = a.cuda(), b.cuda()
    r = ref_chamfer(rl, ad, bd)
    ours = pp.nndistance(ad, bd)
    for x, y, name in zip(ours, r, ["dist1", "dist2", "idx1", "idx2"]):
        assert torch.equal(x, y), "ours vs reference: " + name
    if B * N * M <= 4 * 2500 * 2500:
        o = oracle_mod.chamfer_fwd(np32(a), np32(b))
        for x, y, name in zip(o, r, ["dist1", "dist2", "idx1", "idx2"]):
            assert np.array_equal(x, np32(y)), "oracle vs reference: " + name


def test_chamfer_backward_three_way(ref, pp, oracle_mod):
    rl, _ = ref
    a, b = uniform_cloud(3, 2500, 203), uniform_cloud(3, 2500, 204)
    ad, bd = a.cuda(), b.cuda()
    d1, d2, i1, i2 = ref_chamfer(rl, ad, bd)
    gd1, gd2 = torch.rand(3, 2500, device="cuda"), torch.rand(3, 2500, device="cuda")
    g1, g2 = torch.zeros_like(ad), torch.zeros_like(bd)
    torch.cuda.synchronize()
    rl.nmdistance_backward(ad, bd, g1, g2, gd1, gd2, i1, i2)
    torch.cuda.synchronize()
    from pytorch_points_b200._ext import losses
    h1, h2 = torch.empty_like(ad), torch.empty_like(bd)
    losses.nmdistance_backward(ad, bd, h1, h2, gd1, gd2, i1, i2)
    o1, o2 = oracle_mod.chamfer_bwd(np32(a), np32(b), np32(gd1), np32(gd2), np32(i1), np32(i2))
    for got, want in [(h1, g1), (h2, g2)]:
        scale = want.abs().max().item()
        assert (got - want).abs().max().item() / scale <= 1e-5
    assert np.abs(o1 - np32(g1)).max() / np.abs(o1).max() <= 1e-5


def test_labeled_three_way(ref, pp, oracle_mod):
    rl, _ = ref
    a, b = uniform_cloud(2, 700, 205), uniform_cloud(2, 900, 206)
    g = torch.Generator().manual_seed(207)
    la = torch.randint(0, 4, (2, 700, 1), generator=g); lb = torch.randint(0, 3, (2, 900, 1), generator=g)
    ad, bd = a.cuda(), b.cuda()
    d1 = torch.zeros(2, 700, device="cuda"); d2 = torch.zeros(2, 900, device="cuda")
    i1 = torch.zeros(2, 700, dtype=torch.int32, device="cuda"); i2 = torch.zeros(2, 900, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    rl.labeled_nmdistance_forward(ad, bd, la.cuda().float(), lb.cuda().float(), d1, d2, i1, i2)
    torch.cuda.synchronize()
    ours = pp.labeled_nndistance(ad, bd, la.cuda(), lb.cuda())
    orc = oracle_mod.chamfer_labeled_fwd(np32(a), np32(b), np32(la.float()), np32(lb.float()))
    for x, y, z in zip(ours, (d1, d2, i1, i2), orc):
        assert torch.equal(x, y)
        assert np.array_equal(z, np32(y))


@pytest.mark.parametrize("B,N,m,maker,seed", [(2, 100, 30, uniform_cloud, 0), (2, 511, 64, uniform_cloud, 7),
                                              (2, 5000, 300, sphere_cloud, 3), (2, 4096, 300, None, 1),
                                              (2, 4096, 300, lattice_cloud, 1), (4, 16384, 512, uniform_cloud, 0)])
def test_fps_three_way(ref, pp, oracle_mod, B, N, m, maker, seed):
    _, rs = ref
    x = with_duplicates(uniform_cloud(B, N, 208), 0.3) if maker is None else maker(B, N, 208)
    xd = x.cuda()
    idx = torch.empty(B, m, dtype=torch.int32, device="cuda")
    temp = torch.full((B, N), 1e10, device="cuda")
    torch.cuda.synchronize()
    rs.furthest_sampling(m, seed, xd, temp, idx)
    torch.cuda.synchronize()
    from pytorch_points_b200._ext import sampling
    idx2 = torch.empty_like(idx)
    temp2 = torch.full((B, N), 1e10, device="cuda")
    sampling.furthest_sampling(m, seed, xd, temp2, idx2)
    assert torch.equal(idx2, idx), "ours vs reference idx"
    assert torch.equal(temp2, temp), "ours vs reference temp"
    oi, ot = oracle_mod.fps(np32(x), m, seed=seed, return_temp=True)
    assert np.array_equal(oi, np32(idx)) and np.array_equal(ot, np32(temp)), "oracle vs reference"


@pytest.mark.parametrize("B,N,M,r,ns,maker", [(2, 4096, 512, 0.2, 32, uniform_cloud), (2, 1000, 200, 0.05, 16, uniform_cloud),
                                               (2, 3000, 100, 0.2, 32, sphere_cloud), (2, 777, 33, 0.25, 8, lattice_cloud),
                                               (16, 16384, 1024, 0.2, 32, uniform_cloud)])
def test_ball_query_group_three_way(ref, pp, oracle_mod, B, N, M, r, ns, maker):
    _, rs = ref
    xyz, ctr = maker(B, N, 209), maker(B, M, 210)
    xd, cd = xyz.cuda(), ctr.cuda()
    want = rs.ball_query(cd, xd, r, ns)
    got = pp.ball_query(r, ns, xd, cd)
    assert torch.equal(got, want)
    if B * N * M <= 2 * 4096 * 512:
        assert np.array_equal(oracle_mod.ball_query(r, ns, np32(xyz), np32(ctr)), np32(want))
    feats = xd.transpose(1, 2).contiguous()
    assert torch.equal(pp.grouping_operation(feats, got), rs.group_points(feats, want))
    # the fused QueryAndGroup kernel vs the reference's op sequence on the reference's kernels
    # (network/operations.py:193-205)
    extra = uniform_cloud(B, N, 212, c=5).transpose(1, 2).contiguous().cuda()
    grouped_xyz = rs.group_points(feats, want)
    grouped_xyz -= cd.transpose(1, 2).unsqueeze(-1)
    ref_out = torch.cat([grouped_xyz, rs.group_points(extra, want)], dim=1)
    out, idx = pp.query_and_group(xd, cd, extra, r, ns, True)
    assert torch.equal(idx, want) and torch.equal(out, ref_out)


def test_gather_three_way(ref, pp):
    _, rs = ref
    f = uniform_cloud(3, 500, 211, c=7).transpose(1, 2).contiguous().cuda()
    idx = torch.randint(0, 500, (3, 123), dtype=torch.int32).cuda()
    out = torch.empty(3, 7, 123, device="cuda")
    rs.gather_forward(3, 7, 500, 123, f, idx, out)
    assert torch.equal(pp.gather_points(f, idx), out)


def test_three_nn_three_way(ref, oracle_mod):
    _, rs = ref
    from pytorch_points_b200._ext import sampling
    u, k = uniform_cloud(2, 777, 212), uniform_cloud(2, 1300, 213)
    d = torch.empty(2, 777, 3, device="cuda"); i = torch.empty(2, 777, 3, dtype=torch.int32, device="cuda")
    rs.three_nn_wrapper(2, 777, 1300, u.cuda(), k.cuda(), d, i)
    d2, i2 = sampling.three_nn(u.cuda(), k.cuda())
    assert torch.equal(d2, d) and torch.equal(i2, i)
    od, oi = oracle_mod.three_nn(np32(u), np32(k))
    assert np.array_equal(od, np32(d)) and np.array_equal(oi, np32(i))


def test_three_interpolate_three_way(ref, oracle_mod):
    _, rs = ref
    from pytorch_points_b200._ext import sampling
    u, k = uniform_cloud(2, 600, 214), uniform_cloud(2, 900, 215)
    d, i = sampling.three_nn(u.cuda(), k.cuda())
    w = 1.0 / (torch.sqrt(d) + 1e-8)
    w = (w / w.sum(dim=2, keepdim=True)).contiguous()
    f = uniform_cloud(2, 900, 216, c=7).transpose(1, 2).contiguous().cuda()
    want = torch.empty(2, 7, 600, device="cuda")
    rs.three_interpolate_wrapper(2, 7, 900, 600, f, i, w, want)
    got = torch.empty(2, 7, 600, device="cuda")
    sampling.three_interpolate_wrapper(2, 7, 900, 600, f, i, w, got)
    assert torch.equal(got, want)
    assert np.array_equal(oracle_mod.three_interpolate_fwd(np32(f), np32(i), np32(w)), np32(want))
    go = torch.rand(2, 7, 600, device="cuda")
    g_ref = torch.zeros(2, 7, 900, device="cuda"); g_our = torch.zeros(2, 7, 900, device="cuda")
    rs.three_interpolate_grad_wrapper(2, 7, 600, 900, go, i, w, g_ref)
    sampling.three_interpolate_grad_wrapper(2, 7, 600, 900, go, i, w, g_our)
    assert (g_our - g_ref).abs().max().item() <= 1e-5 * g_ref.abs().max().item()
