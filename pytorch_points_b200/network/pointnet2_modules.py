"""PointNet++ set-abstraction / feature-propagation levels on the B200 operators.

The reference's `network/pointnet2_modules.py:12-153` is torch.nn glue around six kernels per scale;
users who keep that file get this repo's kernels by pointing `pytorch_points._ext` at
`pytorch_points_b200._ext` (tests/test_dropin_gpu.py runs the reference's file that
way).  This module is NOT a transcription of it: it is the stage as this repo would build it --

  * `SAStage` runs the sampling + grouping half of a level as ONE sequence on one stream:
    `pp_fps_gather` (FPS with the gather fused), one `pp_channels_to_points` staging of the features,
    then one `pp_query_group_fwd_pm` per scale, all scales sharing the sampled centres and the staged
    features, and -- for fixed shapes without autograd (inference, or
    the geometry half of a frozen encoder) -- replays that sequence as a single CUDA graph;
  * pooling is a reduction over the sample axis (`max` / `mean`), not a 2-D pooling window;
  * the caller's `mlp` lists are left untouched (the reference adds 3 to `mlp[0]` in place).

`PointnetSAModuleMSG`, `PointnetSAModule`, `PointnetFPModule` keep the reference's constructor
keywords, sub-module names (`groupers`, `mlps`, `mlp`) and return values, so checkpoints and call
sites carry over.
"""
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import pointnet2_utils
from .geo_operations import furthest_point_sample
from .layers import SharedMLP
from .operations import QueryAndGroup, stage_features

# max over the sample axis with the gradient routed to the first maximum, like the reference's max_pool2d
_POOL = {"max_pool": lambda t: t.max(dim=-1).values, "avg_pool": lambda t: t.mean(dim=-1)}


class SAStage:
    """Sampling + multi-scale grouping of one set-abstraction level.

    `stage(xyz, features, new_xyz)` -> `(new_xyz, [grouped_k for every scale])`, with
    `grouped_k` of shape (B, 3 + C, npoint, nsample_k) (or (B, C, ...) without `use_xyz`).
    With `graph=True` and no gradient required the FPS + grouping launches are captured once per
    input shape and replayed as one CUDA graph (static input/output buffers, copy-in / copy-out)."""

    def __init__(self, npoint: Optional[int], groupers: Sequence[nn.Module], graph: bool = False):
        self.npoint, self.groupers, self.graph = npoint, groupers, graph
        self._graphs = {}

    def _run(self, xyz, features, new_xyz):
        if new_xyz is None and self.npoint is not None:
            new_xyz = furthest_point_sample(xyz, self.npoint, NCHW=False)[1]
        # the features are staged point-major once and every scale gathers from the copy (full lines per ball
        # member instead of a sector per value); a single scale with few channels does not repay the copy
        staged = None
        if features is not None and features.is_cuda and self.npoint is not None and (
                len(self.groupers) > 1 or features.shape[1] >= 32):
            staged = stage_features(features)
        if staged is None:
            return new_xyz, [g(xyz, new_xyz, features) for g in self.groupers]
        return new_xyz, [g(xyz, new_xyz, features, staged) for g in self.groupers]

    def __call__(self, xyz, features=None, new_xyz=None):
        needs_grad = torch.is_grad_enabled() and (xyz.requires_grad or (features is not None and features.requires_grad))
        if not self.graph or needs_grad or new_xyz is not None or not xyz.is_cuda:
            return self._run(xyz, features, new_xyz)
        key = (xyz.device, tuple(xyz.shape), None if features is None else tuple(features.shape))
        entry = self._graphs.get(key)
        if entry is None:
            sx = xyz.detach().clone().contiguous()
            sf = None if features is None else features.detach().clone().contiguous()
            side = torch.cuda.Stream(device=xyz.device)
            side.wait_stream(torch.cuda.current_stream(xyz.device))
            with torch.cuda.stream(side), torch.no_grad():
                self._run(sx, sf, None)  # warm-up outside capture (lazy module / library initialisation)
            torch.cuda.current_stream(xyz.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(g):
                out = self._run(sx, sf, None)
            entry = self._graphs[key] = (g, sx, sf, out)
        g, sx, sf, out = entry
        sx.copy_(xyz)
        if sf is not None:
            sf.copy_(features)
        g.replay()
        centres, grouped = out
        return (None if centres is None else centres.clone()), [t.clone() for t in grouped]


class PointnetSAModuleMSG(nn.Module):
    """Set-abstraction level with multi-scale grouping: (xyz (B,N,3), features (B,C,N)) ->
    (new_xyz (B,npoint,3), new_features (B, sum_k mlps[k][-1], npoint))."""

    def __init__(self, *, npoint: Optional[int], radii: List[float], nsamples: List[int], mlps: List[List[int]],
                 bn: bool = True, use_xyz: bool = True, pool_method: str = "max_pool", normalization="batch",
                 graph: bool = False):
        super().__init__()
        if not (len(radii) == len(nsamples) == len(mlps)):
            raise ValueError("radii, nsamples and mlps must have one entry per scale")
        if pool_method not in _POOL:
            raise NotImplementedError(pool_method)
        self.npoint, self.pool_method = npoint, pool_method
        self.groupers = nn.ModuleList(
            QueryAndGroup(r, k, use_xyz=use_xyz) if npoint is not None else pointnet2_utils.GroupAll(use_xyz)
            for r, k in zip(radii, nsamples))
        self.mlps = nn.ModuleList(
            SharedMLP([spec[0] + (3 if use_xyz else 0)] + list(spec[1:]), normalization=normalization, activation="relu")
            for spec in mlps)
        self._stage = SAStage(npoint, self.groupers, graph=graph)

    def forward(self, xyz: torch.Tensor, features: torch.Tensor = None, new_xyz=None):
        new_xyz, grouped = self._stage(xyz, features, new_xyz)
        pool = _POOL[self.pool_method]
        return new_xyz, torch.cat([pool(mlp(g)) for mlp, g in zip(self.mlps, grouped)], dim=1)


class PointnetSAModule(PointnetSAModuleMSG):
    """Single-scale set-abstraction level (npoint=None: one group holding the whole cloud)."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None,
                 bn: bool = True, use_xyz: bool = True, pool_method: str = "max_pool", normalization="batch",
                 graph: bool = False):
        super().__init__(npoint=npoint, radii=[radius], nsamples=[nsample], mlps=[mlp], bn=bn, use_xyz=use_xyz,
                         pool_method=pool_method, normalization=normalization, graph=graph)


class PointnetFPModule(nn.Module):
    """Feature propagation: inverse-distance interpolation of `known_feats` (B,C2,m) at the
    `unknown` positions (B,n,3) from their three nearest `known` points (B,m,3), concatenated with
    `unknow_feats` (B,C1,n), through a shared MLP -> (B, mlp[-1], n)."""

    def __init__(self, *, mlp: List[int], normalization: str = "batch"):
        super().__init__()
        self.mlp = SharedMLP(list(mlp), normalization=normalization, activation="relu")

    def forward(self, unknown, known, unknow_feats, known_feats):
        if known is None:  # nothing to interpolate from: broadcast the single global descriptor
            spread = known_feats.expand(known_feats.shape[0], known_feats.shape[1], unknown.shape[1])
        else:
            spread = pointnet2_utils.propagate_features(unknown, known, known_feats)
        stacked = spread if unknow_feats is None else torch.cat([spread, unknow_feats], dim=1)
        return self.mlp(stacked.unsqueeze(-1)).squeeze(-1)
