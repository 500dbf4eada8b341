"""The layer wrappers the set-abstraction / feature-propagation modules need, and the two
edge-convolution layers that call the KNN / FPS / gather operators.

`Conv2d` and `SharedMLP` mirror `pytorch_points.network.layers.Conv2d` / `SharedMLP`
(network/layers.py:9-21,136-183): same constructor arguments, same sub-module names
(`conv`, `norm`, `act`, `layer{i}`) so state dicts interchange.  They are plain torch.nn
plumbing around the hot-path operators; the rest of the reference's layer zoo is out of scope
(SURVEY.md section 8)."""
from typing import List

import torch.nn as nn


class Conv2d(nn.Module):
    """2-D convolution followed by optional normalization and activation."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True,
                 activation=None, normalization=None, momentum=0.01, conv_params={}):
        super().__init__()
        self.activation = activation
        self.normalization = normalization
        bias = not normalization and bias
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                              bias=bias, **conv_params)
        if normalization is not None:
            if normalization == "batch":
                self.norm = nn.BatchNorm2d(out_channels, affine=True, eps=0.001, momentum=momentum)
            elif normalization == "instance":
                self.norm = nn.InstanceNorm2d(out_channels, affine=True, eps=0.001, momentum=momentum)
            else:
                raise ValueError("only \"batch/instance\" normalization permitted.")
        if activation is not None:
            if activation == "relu":
                self.act = nn.ReLU()
            elif activation == "elu":
                self.act = nn.ELU(alpha=1.0)
            elif activation == "lrelu":
                self.act = nn.LeakyReLU(0.1)
            elif activation == "tanh":
                self.act = nn.Tanh()
            else:
                raise ValueError("only \"relu/elu/lrelu/tanh\" implemented")

    def forward(self, x, epoch=None):
        x = self.conv(x)
        if self.normalization is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.act(x)
        return x


class SharedMLP(nn.Sequential):
    """A stack of 1x1 `Conv2d` blocks applied to every (point, sample) position."""

    def __init__(self, args: List[int], activation: str = None, normalization: str = None, **kwargs):
        super().__init__()
        for i in range(len(args) - 1):
            self.add_module("layer{}".format(i),
                            Conv2d(args[i], args[i + 1], 1, normalization=normalization, activation=activation))


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
        import torch
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
        import torch
        for i, mlp in enumerate(self.mlps):
            if i == 0:
                y, idx = self.get_local_graph(x, k=self.k, idx=idx)
                x = x.unsqueeze(-1).repeat(1, 1, 1, self.k)
                y = torch.cat([nn.functional.relu_(mlp(y)), x], dim=1)
            elif i == (self.n - 1):
                y = torch.cat([mlp(y), y], dim=1)
            else:
                y = torch.cat([nn.functional.relu_(mlp(y)), y], dim=1)
        y, _ = torch.max(y, dim=-1)
        return y, idx


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
        import torch
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
        import torch
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
        for i, mlp in enumerate(self.mlps):
            if i == 0:
                y, _ = self.get_local_graph(sampled_x, x, k=self.k)
                centre = sampled_x.unsqueeze(-1).expand(-1, -1, -1, self.k)
                y = torch.cat([nn.functional.relu_(mlp(y)), centre], dim=1)
            elif i == (self.n - 1):
                y = torch.cat([mlp(y), y], dim=1)
            else:
                y = torch.cat([nn.functional.relu_(mlp(y)), y], dim=1)
        y, _ = torch.max(y, dim=-1)
        return y, sampled_xyz, sampled_idx
