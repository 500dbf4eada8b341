"""Parity at BASELINE.json's full sizes (VERDICT r01 "missing" #2 / weak #1) and of the optional Chamfer
sweep path (chamfer_variant 51) against the default exact kernel.

  * config 4: group_knn k=16, B=4, N=131072 -- 4096 sampled queries per cloud against the CPU oracle
    (their rows of the self-KNN must agree bit for bit), and the pruned sweep == the unpruned sweep on
    every row;
  * config 5 per-rank job: Chamfer B=32, N=M=8192, the FULL batch against the reference's own kernels
    (oracle/_ref, 7.5 ms) -- forward bit for bit, backward to 1e-5;
  * chamfer_variant 51 (approximate sweep + exact resolution): dist / idx bit-equal to variant 1 (exact FFMA kernel) on
    plain, tied, lattice, far-from-origin, clustered and degenerate inputs, fused gradients close.
"""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import lattice_cloud, np32, sphere_cloud, uniform_cloud, with_duplicates

pytestmark = pytest.mark.gpu

REF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


@pytest.fixture(scope="module")
def pp():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pytorch_points_b200 import network
    return network


def test_knn_config4_size_sampled_oracle_and_pruning(pp, oracle_mod):
    from pytorch_points_b200 import _C
    from pytorch_points_b200._ext import sampling
    B, N, k, S = 4, 131072, 16, 4096
    pts = uniform_cloud(B, N, 404)
    pd = pts.cuda()
    dist, idx = sampling.knn(k, pd, pd)
    # 4096 sampled queries of every cloud against the oracle (B*S*N = 2.1e9 pairs on the host cores)
    g = torch.Generator().manual_seed(405)
    rows = torch.stack([torch.randperm(N, generator=g)[:S] for _ in range(B)])
    q = torch.gather(pts, 1, rows.unsqueeze(-1).expand(B, S, 3)).contiguous()
    od, oi = oracle_mod.knn(k, np32(q), np32(pts))
    sel = rows.cuda().unsqueeze(-1).expand(B, S, k)
    assert np.array_equal(np32(torch.gather(idx, 1, sel)), oi), "config-4 KNN: sampled rows, indices"
    assert np.array_equal(np32(torch.gather(dist, 1, sel)), od), "config-4 KNN: sampled rows, distances"
    assert int((idx[:, :, 0] != torch.arange(N, device="cuda").view(1, N)).sum()) == 0  # self is the nearest
    assert bool((dist[:, :, 1:] >= dist[:, :, :-1]).all())  # ascending
    # exact pruning: the sweep that skips tiles returns what the sweep that visits all of them returns
    _C.set_option("knn_prune", 0)
    try:
        dist_all, idx_all = sampling.knn(k, pd, pd)
    finally:
        _C.set_option("knn_prune", 1)
    assert torch.equal(idx, idx_all) and torch.equal(dist, dist_all), "pruned != unpruned"
    # the tensor-core path (knn_tc.cu) returns the same rows
    _C.set_option("knn_tc", 1)
    try:
        dist_tc, idx_tc = sampling.knn(k, pd, pd)
    finally:
        _C.set_option("knn_tc", -1)
    assert torch.equal(idx, idx_tc) and torch.equal(dist, dist_tc), "tensor-core path != ordered sweep"


def test_chamfer_config5_rank_job_vs_reference_kernels(pp):
    if not os.path.exists(os.path.join(REF_DIR, "ref_losses.so")):
        pytest.skip("oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)")
    sys.path.insert(0, REF_DIR)
    import ref_losses
    B, N = 32, 8192
    a, b = uniform_cloud(B, N, 501).cuda(), uniform_cloud(B, N, 502).cuda()
    d1 = torch.zeros(B, N, device="cuda"); d2 = torch.zeros(B, N, device="cuda")
    i1 = torch.zeros(B, N, dtype=torch.int32, device="cuda"); i2 = torch.zeros(B, N, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ref_losses.nmdistance_forward(a, b, d1, d2, i1, i2)  # legacy default stream
    gd1, gd2 = torch.rand(B, N, device="cuda"), torch.rand(B, N, device="cuda")
    r1, r2 = torch.zeros_like(a), torch.zeros_like(b)
    ref_losses.nmdistance_backward(a, b, r1, r2, gd1, gd2, i1, i2)
    torch.cuda.synchronize()
    ad, bd = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    o1, o2, j1, j2 = pp.nndistance(ad, bd)
    for x, y, name in [(o1, d1, "dist1"), (o2, d2, "dist2"), (j1, i1, "idx1"), (j2, i2, "idx2")]:
        assert torch.equal(x.detach(), y), "full batch vs reference kernel: " + name
    ((o1 * gd1).sum() + (o2 * gd2).sum()).backward()
    for got, want, name in [(ad.grad, r1, "grad1"), (bd.grad, r2, "grad2")]:
        assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max()), name


def _clustered(B, N, seed):
    g = torch.Generator().manual_seed(seed)
    centres = torch.rand(B, 8, 3, generator=g) * 100.0
    pick = torch.randint(0, 8, (B, N), generator=g)
    return torch.gather(centres, 1, pick.unsqueeze(-1).expand(B, N, 3)) + 1e-3 * torch.rand(B, N, 3, generator=g)


_SWEEP_CASES = {
    "tiny": lambda: (uniform_cloud(1, 1, 91), uniform_cloud(1, 1, 92)),
    "ragged_33x5000": lambda: (uniform_cloud(2, 33, 91), uniform_cloud(2, 5000, 92)),
    "ragged_4500x257": lambda: (uniform_cloud(3, 4500, 91), uniform_cloud(3, 257, 92)),
    "atlasnet_2500": lambda: (uniform_cloud(2, 2500, 91), uniform_cloud(2, 2500, 92)),
    "off_tile_255x257": lambda: (uniform_cloud(2, 255, 91), uniform_cloud(2, 257, 92)),
    "target_8192": lambda: (uniform_cloud(2, 8192, 91), uniform_cloud(2, 8192, 92)),
    "duplicates": lambda: (with_duplicates(uniform_cloud(2, 2500, 5)), with_duplicates(uniform_cloud(2, 3000, 6))),
    "lattice": lambda: (lattice_cloud(2, 1500, 7), lattice_cloud(2, 1300, 8)),
    "same_cloud": lambda: (uniform_cloud(2, 2000, 9), uniform_cloud(2, 2000, 9)),
    "offset_1e3": lambda: (uniform_cloud(2, 3000, 10) + 1000.0, uniform_cloud(2, 3000, 11) + 1000.0),
    "far_apart": lambda: (uniform_cloud(2, 3000, 12) + 50.0, uniform_cloud(2, 3000, 13) - 50.0),
    "clustered": lambda: (_clustered(2, 4000, 14), _clustered(2, 4000, 14) + 1e-4),
    "tiny_scale": lambda: (uniform_cloud(2, 3000, 15) * 1e-4, uniform_cloud(2, 3000, 16) * 1e-4),
    "sphere": lambda: (sphere_cloud(2, 5000, 17), sphere_cloud(2, 5000, 18)),
    "all_equal": lambda: (torch.ones(2, 600, 3), torch.ones(2, 500, 3)),
}


@pytest.mark.parametrize("case", sorted(_SWEEP_CASES))
def test_chamfer_sweep_variant_is_bit_exact(pp, case):
    """The approximate sweep never decides a result: whatever it cannot prove is re-evaluated with the
    reference's exact chain, so dist / idx equal the default kernel's on every input."""
    from pytorch_points_b200 import _C
    from pytorch_points_b200._ext import losses
    a, b = (t.cuda().contiguous() for t in _SWEEP_CASES[case]())
    B, N, M = a.shape[0], a.shape[1], b.shape[1]
    gw = torch.tensor([0.5 / (B * N), 2.0 / (B * M)], device="cuda")

    def run(variant, fused):
        d1 = torch.full((B, N), -7.0, device="cuda"); d2 = torch.full((B, M), -7.0, device="cuda")
        i1 = torch.full((B, N), -7, dtype=torch.int32, device="cuda"); i2 = torch.full((B, M), -7, dtype=torch.int32, device="cuda")
        sums = torch.zeros(2, device="cuda")
        g1, g2 = torch.full_like(a, 3.0), torch.full_like(b, -3.0)
        _C.set_option("chamfer_variant", variant)
        try:
            if fused:
                losses.nmdistance_forward_backward_uniform(a, b, d1, d2, i1, i2, sums, gw, g1, g2)
            else:
                losses.nmdistance_forward(a, b, d1, d2, i1, i2, sums=sums)
                losses.nmdistance_backward_uniform(a, b, g1, g2, gw, i1, i2)
        finally:
            _C.set_option("chamfer_variant", 0)
        torch.cuda.synchronize()
        return d1, d2, i1, i2, sums, g1, g2

    want = run(1, False)
    for fused in (False, True, True):  # twice: the scratch buffer carries no state between calls
        got = run(51, fused)
        for x, y, name in zip(got[:4], want[:4], ["dist1", "dist2", "idx1", "idx2"]):
            assert torch.equal(x, y), "%s (fused=%s)" % (name, fused)
        assert torch.allclose(got[4], want[4], rtol=1e-4)
        for x, y in zip(got[5:], want[5:]):
            assert float((x - y).abs().max()) <= 1e-5 * float(y.abs().max() + 1e-30)


def test_workspace_clean_promise_is_checked_on_request(pp):
    """a caller-owned workspace passed as clean (PP_CHAMFER_WS_CLEAN) skips the per-call
    fill; the debug option chamfer_ws_check verifies the promise and refuses a dirty buffer."""
    from pytorch_points_b200 import _C
    from pytorch_points_b200._ext import losses
    B, N, M = 2, 700, 900
    a, b = uniform_cloud(B, N, 61).cuda(), uniform_cloud(B, M, 62).cuda()
    out = lambda: (torch.empty(B, N, device="cuda"), torch.empty(B, M, device="cuda"),
                   torch.empty(B, N, dtype=torch.int32, device="cuda"), torch.empty(B, M, dtype=torch.int32, device="cuda"))
    want = out(); losses.nmdistance_forward(a, b, *want)
    ws = torch.full((int(_C.lib.pp_chamfer_fwd_workspace_bytes(B, N, M)),), 0xFF, dtype=torch.uint8, device="cuda")
    _C.set_option("chamfer_ws_check", 1)
    try:
        for _ in range(3):  # every call leaves the keys all-ones again
            got = out(); losses.nmdistance_forward(a, b, *got, workspace=ws, workspace_clean=True)
            assert all(torch.equal(x, y) for x, y in zip(got, want))
        ws[12345 % ws.numel()] = 0
        with pytest.raises(RuntimeError, match="not all-ones"):
            losses.nmdistance_forward(a, b, *out(), workspace=ws, workspace_clean=True)
    finally:
        _C.set_option("chamfer_ws_check", 0)
