"""CPU-only checks of the oracle itself: hand-computable cases, tie-breaks, and agreement with
straightforward float64 numpy on generic inputs.  (The oracle is pinned against the reference's
own kernels by tests/golden/*.npz, see test_golden.py.)"""
import numpy as np
import pytest

from helpers import lattice_cloud, np32, uniform_cloud, with_duplicates


def test_chamfer_matches_float64_bruteforce(oracle_mod):
    a = np32(uniform_cloud(2, 300, 1))
    b = np32(uniform_cloud(2, 257, 2))
    d1, d2, i1, i2 = oracle_mod.chamfer_fwd(a, b)
    D = ((a[:, :, None, :].astype(np.float64) - b[:, None, :, :]) ** 2).sum(-1)
    assert np.allclose(d1, D.min(2), rtol=1e-5, atol=1e-7)
    assert np.allclose(d2, D.min(1), rtol=1e-5, atol=1e-7)
    assert (i1 == D.argmin(2)).mean() > 0.99 and (i2 == D.argmin(1)).mean() > 0.99


def test_chamfer_lowest_index_on_ties(oracle_mod):
    a = np.zeros((1, 3, 3), np.float32)
    b = np.ones((1, 600, 3), np.float32)  # all refs equidistant, spans two 512-chunks
    d1, d2, i1, i2 = oracle_mod.chamfer_fwd(a, b)
    assert (i1 == 0).all() and (i2 == 0).all()
    assert np.all(d1 == 3.0) and np.all(d2 == 3.0)


def test_chamfer_generic_dim(oracle_mod):
    a = np32(uniform_cloud(1, 50, 3, c=5))
    b = np32(uniform_cloud(1, 70, 4, c=5))
    d1, _, i1, _ = oracle_mod.chamfer_fwd(a, b)
    D = ((a[:, :, None, :].astype(np.float64) - b[:, None, :, :]) ** 2).sum(-1)
    assert np.allclose(d1, D.min(2), rtol=1e-5)
    assert (i1 == D.argmin(2)).all()


def test_labeled_chamfer(oracle_mod):
    a = np32(uniform_cloud(1, 40, 5))
    b = np32(uniform_cloud(1, 30, 6))
    la = (np.arange(40) % 3).astype(np.float32)[None]
    lb = (np.arange(30) % 2).astype(np.float32)[None]  # label 2 never appears in b
    d1, d2, i1, i2 = oracle_mod.chamfer_labeled_fwd(a, b, la, lb)
    assert (i1[0][la[0] == 2] == -1).all() and (d1[0][la[0] == 2] == 0).all()
    ok = la[0] != 2
    assert (lb[0][i1[0][ok]] == la[0][ok]).all()
    assert (i2 >= 0).all()


def test_chamfer_bwd_matches_autograd_formula(oracle_mod):
    a = np32(uniform_cloud(2, 64, 7))
    b = np32(uniform_cloud(2, 48, 8))
    d1, d2, i1, i2 = oracle_mod.chamfer_fwd(a, b)
    gd1 = np32(uniform_cloud(2, 64, 9, c=1))[..., 0]
    gd2 = np32(uniform_cloud(2, 48, 10, c=1))[..., 0]
    g1, g2 = oracle_mod.chamfer_bwd(a, b, gd1, gd2, i1, i2)
    e1 = np.zeros_like(a, dtype=np.float64)
    e2 = np.zeros_like(b, dtype=np.float64)
    for bb in range(2):
        for j in range(64):
            v = 2 * gd1[bb, j] * (a[bb, j].astype(np.float64) - b[bb, i1[bb, j]])
            e1[bb, j] += v
            e2[bb, i1[bb, j]] -= v
        for j in range(48):
            v = 2 * gd2[bb, j] * (b[bb, j].astype(np.float64) - a[bb, i2[bb, j]])
            e2[bb, j] += v
            e1[bb, i2[bb, j]] -= v
    assert np.allclose(g1, e1, rtol=1e-5, atol=1e-6) and np.allclose(g2, e2, rtol=1e-5, atol=1e-6)


def test_fps_block_size_rule(oracle_mod):
    for n, bs in [(1, 1), (2, 2), (3, 2), (511, 256), (512, 512), (513, 512), (16384, 512), (100000, 512)]:
        assert oracle_mod.fps_block_size(n) == bs


def test_fps_basic_and_seed(oracle_mod):
    x = np32(uniform_cloud(2, 700, 11))
    idx = oracle_mod.fps(x, 32, seed=5)
    assert (idx[:, 0] == 5).all()
    for b in range(2):
        assert len(set(idx[b].tolist())) == 32
    # greedy property: sample j maximises the distance to the already chosen set
    for b in range(2):
        chosen = [idx[b, 0]]
        for j in range(1, 8):
            d = ((x[b][:, None, :].astype(np.float64) - x[b][chosen][None]) ** 2).sum(-1).min(1)
            assert np.isclose(d[idx[b, j]], d.max(), rtol=1e-6)
            chosen.append(idx[b, j])


def test_fps_tie_key_is_k_mod_bs_then_k(oracle_mod):
    # 1024 points: point 0 at origin, points 600 and 513 exactly tied at the max distance.
    x = np.zeros((1, 1024, 3), np.float32)
    x[0, 600] = [2, 0, 0]   # 600 % 512 == 88
    x[0, 513] = [0, 2, 0]   # 513 % 512 == 1   -> wins although 513 < 600 is NOT why: slot 1 < slot 88
    x[0, 90] = [0, 0, 2]    # 90 % 512 == 90   -> lowest index but highest slot: loses
    idx = oracle_mod.fps(x, 2, seed=0)
    assert idx[0, 1] == 513


def test_fps_all_duplicates_picks_zero(oracle_mod):
    x = np.ones((1, 100, 3), np.float32)
    idx = oracle_mod.fps(x, 5, seed=3)
    assert idx[0].tolist() == [3, 0, 0, 0, 0]


def test_ball_query_semantics(oracle_mod):
    xyz = np.zeros((1, 10, 3), np.float32)
    xyz[0, :, 0] = np.arange(10) * 0.1
    centres = np.array([[[0.35, 0, 0], [5.0, 0, 0], [0.0, 0, 0]]], np.float32)
    idx = oracle_mod.ball_query(0.16, 4, xyz, centres)
    assert idx[0, 0].tolist() == [2, 3, 4, 5]       # first 4 hits in index order
    assert idx[0, 1].tolist() == [0, 0, 0, 0]       # empty ball -> zeros
    assert idx[0, 2].tolist() == [0, 1, 0, 0]       # padded with the first hit
    # strict <: a point exactly at distance r is outside
    idx = oracle_mod.ball_query(0.5, 2, np.array([[[0.5, 0, 0]]], np.float32), np.zeros((1, 1, 3), np.float32))
    assert idx[0, 0].tolist() == [0, 0]


def test_gather_and_group(oracle_mod):
    f = np32(uniform_cloud(2, 20, 12, c=4)).transpose(0, 2, 1).copy()  # (B,C,N)
    idx = np.array([[3, 3, 19], [0, 7, 7]], np.int32)
    out = oracle_mod.gather_fwd(f, idx)
    assert out.shape == (2, 4, 3) and np.array_equal(out[1, :, 1], f[1, :, 7])
    g = oracle_mod.gather_bwd(np.ones_like(out), idx, 20)
    assert g[0, 0, 3] == 2 and g[0, 0, 19] == 1 and g.sum() == out.size
    gidx = np.array([[[1, 2], [2, 2]], [[0, 0], [5, 6]]], np.int32)
    go = oracle_mod.group_fwd(f, gidx)
    assert go.shape == (2, 4, 2, 2) and np.array_equal(go[1, :, 1, 1], f[1, :, 6])
    gg = oracle_mod.group_bwd(np.ones_like(go), gidx, 20)
    assert gg[0, 0, 2] == 3


def test_knn_order_and_ties(oracle_mod):
    p = np32(lattice_cloud(1, 200, 13, levels=4))
    d, i = oracle_mod.knn(8, p[:, :10], p)
    D = ((p[:, :10, None, :].astype(np.float64) - p[:, None, :, :]) ** 2).sum(-1)[0]
    for q in range(10):
        order = np.lexsort((np.arange(200), D[q]))[:8]   # (distance, index) ascending
        assert i[0, q].tolist() == order.tolist()
        assert np.allclose(d[0, q], D[q][order], rtol=1e-6)


def test_three_nn(oracle_mod):
    u = np32(uniform_cloud(1, 20, 14))
    k = np32(uniform_cloud(1, 50, 15))
    d, i = oracle_mod.three_nn(u, k)
    D = ((u[:, :, None, :].astype(np.float64) - k[:, None, :, :]) ** 2).sum(-1)[0]
    assert (i[0] == np.argsort(D, axis=1, kind="stable")[:, :3]).all()
    assert np.allclose(d[0], np.sort(D, axis=1)[:, :3], rtol=1e-5)


def test_duplicates_helper_creates_ties():
    x = with_duplicates(uniform_cloud(1, 100, 16))
    assert (x[0, 90:] == x[0, :10]).all()
