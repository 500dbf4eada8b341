"""The layer wrappers the set-abstraction / feature-propagation modules need, and the two
edge-convolution layers that call the KNN / FPS / gather operators.

`Conv2d` and `SharedMLP` mirror `pytorch_points.network.layers.Conv2d` / `SharedMLP`
(network/layers.py:9-21,136-183): same constructor arguments, same sub-module names
(`conv`, `norm`, `act`, `layer{i}`) so state dicts interchange.  They are plain torch.nn
plumbing around the hot-path operators; the rest of the reference's layer zoo is out of scope
(SURVEY.md section 8)."""
from typing import List, Optional

import torch
import torch.nn as nn

# what the `normalization` / `activation` keywords of the reference's layers select (network/layers.py:9-21)
_NORMS = {
    "batch": lambda ch, momentum: nn.BatchNorm2d(ch, affine=True, eps=0.001, momentum=momentum),
    "instance": lambda ch, momentum: nn.InstanceNorm2d(ch, affine=True, eps=0.001, momentum=momentum),
}
_ACTIVATIONS = {
    "relu": nn.ReLU,
    "elu": lambda: nn.ELU(alpha=1.0),
    "lrelu": lambda: nn.LeakyReLU(0.1),
    "tanh": nn.Tanh,
}


def _pick(table, key, what):
    if key not in table:
        raise ValueError("%s %r is not one of %s" % (what, key, sorted(table)))
    return table[key]


class Conv2d(nn.Module):
    """Convolution -> optional normalization (`norm`) -> optional activation (`act`); the convolution carries a
    bias only when nothing normalizes its output.  Keywords and sub-module names as in the reference."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True,
                 activation: Optional[str] = None, normalization: Optional[str] = None, momentum=0.01, conv_params=None):
        super().__init__()
        self.activation, self.normalization = activation, normalization
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                              bias=bool(bias) and not normalization, **(conv_params or {}))
        self.norm = None if normalization is None else _pick(_NORMS, normalization, "normalization")(out_channels, momentum)
        self.act = None if activation is None else _pick(_ACTIVATIONS, activation, "activation")()

    def forward(self, x, epoch=None):
        for stage in (self.conv, self.norm, self.act):
            if stage is not None:
                x = stage(x)
        return x


class SharedMLP(nn.Sequential):
    """1x1 `Conv2d` blocks `layer0`, `layer1`, ... with widths `args[0] -> args[1] -> ...`, applied to every
    (point, sample) position."""

    def __init__(self, args: List[int], activation: str = None, normalization: str = None, **kwargs):
        super().__init__()
        for i, (c_in, c_out) in enumerate(zip(args[:-1], args[1:])):
            self.add_module("layer%d" % i, Conv2d(c_in, c_out, 1, normalization=normalization, activation=activation))


def _dense_stack(mlps, edge, carry):
    """The densely connected part shared by both edge convolutions: the first MLP sees the edge features and is
    concatenated with `carry` (the centre features repeated over the k neighbours); every later MLP sees everything
    produced so far and is stacked in front of it; ReLU after all but the last; finally the maximum over the
    neighbours.  -> (B, C', S)"""
    last = len(mlps) - 1
    y = edge
    for i, mlp in enumerate(mlps):
        out = mlp(y)
        if i != last or i == 0:
            out = nn.functional.relu_(out)
        y = torch.cat([out, carry if i == 0 else y], dim=1)
    return y.max(dim=-1)[0]


class DenseEdgeConv(nn.Module):
    """Densely connected edge convolution over a k-NN graph (reference network/layers.py:23-82),
    with the neighbour search on `operations.knn_points` (this repo's KNN kernels) instead of
    pytorch3d -- SURVEY.md next row N4."""

    def __init__(self, in_channels, growth_rate, n, k, **kwargs):
        super().__init__()
        self.growth_rate, self.n, self.k = growth_rate, n, k
        self.mlps = nn.ModuleList([nn.Conv2d(2 * in_channels, growth_rate, 1, bias=True)])
        for _ in range(1, n):
            in_channels += growth_rate
            self.mlps.append(nn.Conv2d(in_channels, growth_rate, 1, bias=True))
        self.out_channels = in_channels + growth_rate

    def get_local_graph(self, x, k, idx=None):
        """x (B, C, N) -> edge features [x_i, x_j - x_i] (B, 2C, N, k) over the k nearest
        neighbours j of i (the point itself excluded), and their indices (B, N, k)."""
        from .operations import knn_points
        pts = x.transpose(1, 2).contiguous()  # (B, N, C)
        if idx is None:
            _, idx, nn_pts = knn_points(pts, pts, K=k + 1, return_nn=True)  # (B, N, k+1[, C])
            idx, nn_pts = idx[:, :, 1:], nn_pts[:, :, 1:, :]
        else:
            B, N, C = pts.shape
            nn_pts = torch.gather(pts.unsqueeze(1).expand(B, N, N, C), 2, idx.long().unsqueeze(-1).expand(B, N, idx.shape[2], C))
        neighbours = nn_pts.permute(0, 3, 1, 2)              # (B, C, N, k)
        centre = x.unsqueeze(-1).expand_as(neighbours)
        return torch.cat([centre, neighbours - centre], dim=1), idx

    def forward(self, x, idx=None):
        """x (B, C, N) -> (features (B, C', N), knn idx (B, N, k))."""
        edge, idx = self.get_local_graph(x, k=self.k, idx=idx)
        carry = x.unsqueeze(-1).expand(-1, -1, -1, self.k)
        return _dense_stack(self.mlps, edge, carry), idx


class SampledDenseEdgeConv(DenseEdgeConv):
    """`DenseEdgeConv` evaluated at `nsample` farthest-point-sampled centres only (reference
    network/layers.py:85-133): FPS on the coordinates -> gather the centres' features -> k nearest
    FEATURE neighbours of each centre among all N points -> the dense edge MLP -> max over k.
    A caller of three hot-path operators (`furthest_point_sample`, `gather_points`, KNN).

    The snapshot's `get_local_graph` unpacks `knn_points(..., return_nn=True)` into two names
    (layers.py:99), which raises for the three-tuple the call returns; this mirror implements what
    the surrounding code evidently intends (idx and neighbours, the nearest entry -- the centre
    itself -- dropped)."""

    def get_local_graph(self, query, x, k, idx=None):
        """query (B, C, S) centres, x (B, C, N) all points -> edge features [q_i, x_j - q_i]
        (B, 2C, S, k) over the k nearest neighbours j of centre i in feature space (the nearest
        one, i itself, excluded) and their indices (B, S, k)."""
        from . import operations as ops
        pts = x.transpose(1, 2).contiguous()  # (B, N, C)
        if idx is None:
            _, idx, nn_pts = ops.knn_points(query.transpose(1, 2).contiguous(), pts, K=k + 1, return_nn=True)
            idx, nn_pts = idx[:, :, 1:], nn_pts[:, :, 1:, :]
        else:
            B, N, C = pts.shape
            S = query.shape[2]
            nn_pts = torch.gather(pts.unsqueeze(1).expand(B, S, N, C), 2,
                                  idx.long().unsqueeze(-1).expand(B, S, idx.shape[2], C))
        neighbours = nn_pts.permute(0, 3, 1, 2)  # (B, C, S, k)
        centre = query.unsqueeze(-1).expand_as(neighbours)
        return torch.cat([centre, neighbours - centre], dim=1), idx

    def forward(self, x, nsample, xyz):
        """x (B, C, N) features, xyz (B, 3, N) coordinates ->
        (features (B, C', nsample), sampled_xyz (B, 3, nsample), sampled_idx (B, nsample))."""
        from . import geo_operations as geo
        from . import operations as ops
        if nsample == 1:
            # one centre: the point nearest to the centroid
            centroid = torch.mean(xyz, dim=-1, keepdim=True)  # (B, 3, 1)
            _, sampled_idx, sampled_xyz = ops.knn_points(centroid.transpose(1, 2).contiguous(),
                                                         xyz.transpose(1, 2).contiguous(), K=1, return_nn=True)
            sampled_xyz = sampled_xyz.squeeze(2).transpose(1, 2).contiguous()  # (B, 1, 1, 3) -> (B, 3, 1)
            sampled_idx = sampled_idx.squeeze(2).int()                          # (B, 1, 1) -> (B, 1)
        else:
            sampled_idx, sampled_xyz = geo.furthest_point_sample(xyz, nsample, NCHW=True)
        sampled_x = ops.gather_points(x.contiguous(), sampled_idx)  # (B, C, nsample)
        edge, _ = self.get_local_graph(sampled_x, x, k=self.k)
        carry = sampled_x.unsqueeze(-1).expand(-1, -1, -1, self.k)
        return _dense_stack(self.mlps, edge, carry), sampled_xyz, sampled_idx
