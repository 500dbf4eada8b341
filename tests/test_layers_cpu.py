"""Host-side logic of the KNN/FPS callers in network/layers.py, without a GPU: the three operators
they call are replaced by plain torch stand-ins (monkeypatch), so only the wiring around them --
shapes, the dropped self-neighbour, the dense MLP stack, the max over k -- is under test."""
import pytest
import torch


def _knn_points(p1, p2, K=1, return_nn=False):
    d = ((p1[:, :, None, :] - p2[:, None, :, :]) ** 2).sum(-1)
    dist, idx = torch.sort(d, dim=-1, stable=True)
    dist, idx = dist[:, :, :K], idx[:, :, :K]
    nn = None
    if return_nn:
        B, M, _ = p1.shape
        nn = torch.gather(p2.unsqueeze(1).expand(B, M, p2.shape[1], p2.shape[2]), 2,
                          idx.unsqueeze(-1).expand(B, M, K, p2.shape[2]))
    return dist, idx, nn


def _fps(xyz, npoint, NCHW=True, seedIdx=0):
    pts = xyz.transpose(1, 2) if NCHW else xyz
    B, N, _ = pts.shape
    idx = torch.zeros(B, npoint, dtype=torch.int32)
    for b in range(B):
        temp = torch.full((N,), 1e10)
        old = seedIdx
        idx[b, 0] = old
        for j in range(1, npoint):
            temp = torch.minimum(temp, ((pts[b] - pts[b, old]) ** 2).sum(-1))
            old = int(torch.argmax(temp))
            idx[b, j] = old
    sampled = torch.gather(pts, 1, idx.long().unsqueeze(-1).expand(B, npoint, 3))
    return idx, (sampled.transpose(1, 2).contiguous() if NCHW else sampled)


def _gather_points(features, idx):
    return torch.gather(features, 2, idx.long().unsqueeze(1).expand(-1, features.shape[1], -1))


@pytest.fixture()
def patched(monkeypatch):
    from pytorch_points_b200.network import geo_operations, layers, operations
    monkeypatch.setattr(operations, "knn_points", _knn_points)
    monkeypatch.setattr(operations, "gather_points", _gather_points)
    monkeypatch.setattr(geo_operations, "furthest_point_sample", _fps)
    return layers


def _dense_stack(mod, y, centre):
    """The DenseNet-style stack of network/layers.py:71-82 written out for one edge tensor."""
    for i, mlp in enumerate(mod.mlps):
        if i == 0:
            y = torch.cat([torch.relu(mlp(y)), centre], dim=1)
        elif i == mod.n - 1:
            y = torch.cat([mlp(y), y], dim=1)
        else:
            y = torch.cat([torch.relu(mlp(y)), y], dim=1)
    return y.max(dim=-1)[0]


@pytest.mark.parametrize("nsample", [7, 1])
def test_sampled_dense_edge_conv_wiring(patched, nsample):
    torch.manual_seed(0)
    B, C, N, k = 2, 4, 40, 5
    mod = patched.SampledDenseEdgeConv(C, growth_rate=6, n=3, k=k)
    x, xyz = torch.rand(B, C, N), torch.rand(B, 3, N)
    y, sxyz, sidx = mod(x, nsample, xyz)
    assert y.shape == (B, mod.out_channels, nsample) and sxyz.shape == (B, 3, nsample) and sidx.shape == (B, nsample)
    # centres: FPS order (or the point nearest the centroid), coordinates gathered from xyz
    if nsample == 1:
        want = ((xyz - xyz.mean(-1, keepdim=True)) ** 2).sum(1).argmin(-1, keepdim=True)
    else:
        want = _fps(xyz, nsample)[0].long()
    assert torch.equal(sidx.long(), want)
    assert torch.equal(sxyz, torch.gather(xyz, 2, want.unsqueeze(1).expand(B, 3, nsample)))
    # features: brute-force edge tensor per centre, the centre itself (nearest, distance 0) dropped
    feats = x.transpose(1, 2)
    edges = torch.empty(B, 2 * C, nsample, k)
    for b in range(B):
        for s in range(nsample):
            c = feats[b, want[b, s]]
            order = torch.sort(((feats[b] - c) ** 2).sum(-1), stable=True)[1][1:k + 1]
            assert int(want[b, s]) not in order.tolist()
            edges[b, :C, s] = c[:, None]
            edges[b, C:, s] = (feats[b, order] - c).t()
    centre = _gather_points(x, want).unsqueeze(-1).expand(-1, -1, -1, k)
    assert torch.allclose(y, _dense_stack(mod, edges, centre), atol=1e-6)
    # a given neighbour index list is honoured
    e2, idx2 = mod.get_local_graph(_gather_points(x, want), x, k, idx=torch.zeros(B, nsample, k, dtype=torch.long))
    assert torch.equal(e2[:, C:], (x[:, :, :1].unsqueeze(-1) - _gather_points(x, want)[..., None]).expand(-1, -1, -1, k))


def test_dense_edge_conv_wiring(patched):
    torch.manual_seed(1)
    B, C, N, k = 2, 3, 30, 4
    mod = patched.DenseEdgeConv(C, growth_rate=5, n=2, k=k)
    x = torch.rand(B, C, N)
    y, idx = mod(x)
    assert y.shape == (B, mod.out_channels, N) and idx.shape == (B, N, k)
    feats = x.transpose(1, 2)
    for b in range(B):
        for i in range(0, N, 7):
            order = torch.sort(((feats[b] - feats[b, i]) ** 2).sum(-1), stable=True)[1][1:k + 1]
            assert idx[b, i].tolist() == order.tolist()
    edges, _ = mod.get_local_graph(x, k)
    assert torch.allclose(y, _dense_stack(mod, edges, x.unsqueeze(-1).expand(-1, -1, -1, k)), atol=1e-6)
