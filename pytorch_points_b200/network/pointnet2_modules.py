"""PointNet++ set-abstraction and feature-propagation modules over the B200 operators.

Drop-in for `pytorch_points.network.pointnet2_modules` (network/pointnet2_modules.py:12-153):
same class names, keyword arguments, sub-module names and return values.  The sampling and
grouping work of a set-abstraction level is three kernels here -- FPS with the gather fused
(csrc/sampling.cu), one fused ball-query + grouping kernel per scale (csrc/sa_group.cu) -- where
the reference launches FPS, gather, and six kernels per scale; the shared MLP and pooling stay
torch.nn (out of the hot path)."""
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointnet2_utils
from .geo_operations import furthest_point_sample
from .layers import SharedMLP
from .operations import QueryAndGroup


class _PointnetSAModuleBase(nn.Module):

    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None
        self.pool_method = 'max_pool'

    def forward(self, xyz: torch.Tensor, features: torch.Tensor = None, new_xyz=None):
        """xyz (B, N, 3), features (B, C, N) -> (new_xyz (B, npoint, 3),
        new_features (B, sum_k mlps[k][-1], npoint))."""
        new_features_list = []
        if new_xyz is None:
            new_xyz = furthest_point_sample(xyz, self.npoint, NCHW=False)[1] if self.npoint is not None else None

        for i in range(len(self.groupers)):
            new_features = self.groupers[i](xyz, new_xyz, features)  # (B, C, npoint, nsample)
            new_features = self.mlps[i](new_features)  # (B, mlp[-1], npoint, nsample)
            if self.pool_method == 'max_pool':
                new_features = F.max_pool2d(new_features, kernel_size=[1, new_features.size(3)])
            elif self.pool_method == 'avg_pool':
                new_features = F.avg_pool2d(new_features, kernel_size=[1, new_features.size(3)])
            else:
                raise NotImplementedError
            new_features_list.append(new_features.squeeze(-1))  # (B, mlp[-1], npoint)

        return new_xyz, torch.cat(new_features_list, dim=1)


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    """Set-abstraction level with multi-scale grouping."""

    def __init__(self, *, npoint: int, radii: List[float], nsamples: List[int], mlps: List[List[int]],
                 bn: bool = True, use_xyz: bool = True, pool_method='max_pool', normalization="batch"):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        for i in range(len(radii)):
            self.groupers.append(QueryAndGroup(radii[i], nsamples[i], use_xyz=use_xyz)
                                 if npoint is not None else pointnet2_utils.GroupAll(use_xyz))
            mlp_spec = mlps[i]
            if use_xyz:
                mlp_spec[0] += 3  # in place, as the reference does (pointnet2_modules.py:88)
            self.mlps.append(SharedMLP(mlp_spec, normalization=normalization, activation="relu"))
        self.pool_method = pool_method


class PointnetSAModule(PointnetSAModuleMSG):
    """Set-abstraction level with a single scale."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None,
                 bn: bool = True, use_xyz: bool = True, pool_method='max_pool', normalization="batch"):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn,
                         use_xyz=use_xyz, pool_method=pool_method, normalization=normalization)


class PointnetFPModule(nn.Module):
    """Propagates the features of one point set to another (three_nn + three_interpolate)."""

    def __init__(self, *, mlp: List[int], normalization: str = "batch"):
        super().__init__()
        self.mlp = SharedMLP(mlp, normalization=normalization, activation="relu")

    def forward(self, unknown, known, unknow_feats, known_feats):
        """unknown (B, n, 3), known (B, m, 3), unknow_feats (B, C1, n), known_feats (B, C2, m)
        -> (B, mlp[-1], n)."""
        if known is not None:
            dist, idx = pointnet2_utils.three_nn(unknown, known)
            dist_recip = 1.0 / (dist + 1e-8)
            norm = torch.sum(dist_recip, dim=2, keepdim=True)
            weight = dist_recip / norm
            interpolated_feats = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        else:
            interpolated_feats = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))

        if unknow_feats is not None:
            new_features = torch.cat([interpolated_feats, unknow_feats], dim=1)  # (B, C2 + C1, n)
        else:
            new_features = interpolated_feats
        return self.mlp(new_features.unsqueeze(-1)).squeeze(-1)
