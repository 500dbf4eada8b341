"""Golden vectors captured from the REFERENCE's own CUDA kernels on a B200
(tests/golden/make_golden.py ran oracle/_ref/*.so, the unmodified reference sources built for
sm_100a).  CPU part: they pin the oracle.  GPU part: our kernels reproduce them.

Bar: indices and distances bit-exact; gradients (atomic summation order) within 1e-5 relative."""
import glob
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, prefix + "*.npz")))


def rel_err(got, want):
    return np.abs(np.asarray(got, np.float64) - want).max() / max(np.abs(want).max(), 1e-30)


def test_fixture_inventory():
    assert len(names("chamfer_")) >= 5 and len(names("fps_")) >= 6 and len(names("bq_")) >= 5
    assert os.path.exists(os.path.join(GOLD, "three_nn.npz"))


# ------------------------------------------------------------------ oracle vs reference (CPU)
@pytest.mark.parametrize("name", [n for n in names("chamfer_") if n != "chamfer_labeled"])
def test_oracle_chamfer_matches_reference(oracle_mod, name):
    g = load(name)
    d1, d2, i1, i2 = oracle_mod.chamfer_fwd(g["xyz1"], g["xyz2"])
    assert np.array_equal(i1, g["idx1"]) and np.array_equal(i2, g["idx2"])
    assert np.array_equal(d1.view(np.uint32), g["dist1"].view(np.uint32))
    assert np.array_equal(d2.view(np.uint32), g["dist2"].view(np.uint32))
    g1, g2 = oracle_mod.chamfer_bwd(g["xyz1"], g["xyz2"], g["gd1"], g["gd2"], g["idx1"], g["idx2"])
    assert rel_err(g1, g["g1"]) <= 1e-5 and rel_err(g2, g["g2"]) <= 1e-5


def test_oracle_labeled_matches_reference(oracle_mod):
    g = load("chamfer_labeled")
    d1, d2, i1, i2 = oracle_mod.chamfer_labeled_fwd(g["xyz1"], g["xyz2"], g["label1"], g["label2"])
    assert np.array_equal(i1, g["idx1"]) and np.array_equal(i2, g["idx2"])
    assert np.array_equal(d1, g["dist1"]) and np.array_equal(d2, g["dist2"])
    assert (g["idx1"] == -1).any()


@pytest.mark.parametrize("name", names("fps_"))
def test_oracle_fps_matches_reference(oracle_mod, name):
    g = load(name)
    idx, temp = oracle_mod.fps(g["xyz"], int(g["m"]), seed=int(g["seed"]), return_temp=True)
    assert np.array_equal(idx, g["idx"])
    assert np.array_equal(temp.view(np.uint32), g["temp"].view(np.uint32))
    feats = np.ascontiguousarray(g["xyz"].transpose(0, 2, 1))
    assert np.array_equal(oracle_mod.gather_fwd(feats, g["idx"]), g["gathered"])
    assert rel_err(oracle_mod.gather_bwd(g["grad_out"], g["idx"], g["xyz"].shape[1]), g["grad_in"]) <= 1e-5


@pytest.mark.parametrize("name", names("bq_"))
def test_oracle_ball_query_matches_reference(oracle_mod, name):
    g = load(name)
    idx = oracle_mod.ball_query(float(g["radius"]), int(g["nsample"]), g["xyz"], g["new_xyz"])
    assert np.array_equal(idx, g["idx"])
    feats = np.ascontiguousarray(g["xyz"].transpose(0, 2, 1))
    assert np.array_equal(oracle_mod.group_fwd(feats, g["idx"]), g["grouped"])
    assert rel_err(oracle_mod.group_bwd(np.ones_like(g["grouped"]), g["idx"], g["xyz"].shape[1]), g["group_grad"]) <= 1e-5


def test_oracle_three_nn_matches_reference(oracle_mod):
    g = load("three_nn")
    d, i = oracle_mod.three_nn(g["unknown"], g["known"])
    assert np.array_equal(i, g["idx"]) and np.array_equal(d, g["dist2"])


def test_oracle_three_interpolate_matches_reference(oracle_mod):
    g = load("three_interpolate")
    out = oracle_mod.three_interpolate_fwd(g["features"], g["idx"], g["weight"])
    assert np.array_equal(out.view(np.uint32), g["out"].view(np.uint32))
    gin = oracle_mod.three_interpolate_bwd(g["grad_out"], g["idx"], g["weight"], g["features"].shape[2])
    assert rel_err(gin, g["grad_in"]) <= 1e-5


# ------------------------------------------------------------------ our kernels vs reference (GPU)
@pytest.fixture(scope="module")
def pp():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pytorch_points_b200 import network
    return network


def cu(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n in names("chamfer_") if n != "chamfer_labeled"])
def test_kernels_chamfer_match_reference(pp, name):
    import torch
    g = load(name)
    a, b = cu(g["xyz1"]).requires_grad_(True), cu(g["xyz2"]).requires_grad_(True)
    d1, d2, i1, i2 = pp.nndistance(a, b)
    assert np.array_equal(i1.cpu().numpy(), g["idx1"]) and np.array_equal(i2.cpu().numpy(), g["idx2"])
    assert np.array_equal(d1.detach().cpu().numpy(), g["dist1"]) and np.array_equal(d2.detach().cpu().numpy(), g["dist2"])
    torch.autograd.backward([d1, d2], [cu(g["gd1"]), cu(g["gd2"])])
    assert rel_err(a.grad.cpu().numpy(), g["g1"]) <= 1e-5 and rel_err(b.grad.cpu().numpy(), g["g2"]) <= 1e-5


@pytest.mark.gpu
def test_kernels_labeled_match_reference(pp):
    g = load("chamfer_labeled")
    d1, d2, i1, i2 = pp.labeled_nndistance(cu(g["xyz1"]), cu(g["xyz2"]), cu(g["label1"]), cu(g["label2"]))
    assert np.array_equal(i1.cpu().numpy(), g["idx1"]) and np.array_equal(i2.cpu().numpy(), g["idx2"])
    assert np.array_equal(d1.cpu().numpy(), g["dist1"]) and np.array_equal(d2.cpu().numpy(), g["dist2"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", names("fps_"))
def test_kernels_fps_match_reference(pp, name):
    g = load(name)
    idx, pts = pp.furthest_point_sample(cu(g["xyz"]), int(g["m"]), NCHW=False, seedIdx=int(g["seed"]))
    assert np.array_equal(idx.cpu().numpy(), g["idx"])
    assert np.array_equal(pts.cpu().numpy().transpose(0, 2, 1), g["gathered"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", names("bq_"))
def test_kernels_ball_query_match_reference(pp, name):
    g = load(name)
    idx = pp.ball_query(float(g["radius"]), int(g["nsample"]), cu(g["xyz"]), cu(g["new_xyz"]))
    assert np.array_equal(idx.cpu().numpy(), g["idx"])
    feats = cu(g["xyz"].transpose(0, 2, 1))
    assert np.array_equal(pp.grouping_operation(feats, idx).cpu().numpy(), g["grouped"])


@pytest.mark.gpu
def test_kernels_three_nn_interpolate_match_reference(pp):
    import torch
    g3 = load("three_nn")
    dist, idx = pp.three_nn(cu(g3["unknown"]), cu(g3["known"]))
    assert np.array_equal(idx.cpu().numpy(), g3["idx"])
    assert np.array_equal((dist * dist).cpu().numpy().round(6).shape, g3["dist2"].shape)
    assert np.array_equal(dist.cpu().numpy(), np.sqrt(g3["dist2"]))
    g = load("three_interpolate")
    f = cu(g["features"]).requires_grad_(True)
    out = pp.three_interpolate(f, cu(g["idx"]), cu(g["weight"]))
    assert np.array_equal(out.detach().cpu().numpy(), g["out"])
    out.backward(cu(g["grad_out"]))
    assert rel_err(f.grad.cpu().numpy(), g["grad_in"]) <= 1e-5
