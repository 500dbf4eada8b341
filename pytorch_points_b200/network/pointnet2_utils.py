"""PointNet++ feature-propagation helpers.

Same public surface as `pytorch_points.network.pointnet2_utils` (`ThreeNN` / `three_nn`, `ThreeInterpolate` /
`three_interpolate`, `GroupAll`; network/pointnet2_utils.py:11-88,127-150) -- SURVEY.md "next" row N3 -- on this
repo's kernels, plus `propagate_features`, the whole interpolation step of a feature-propagation level in one
call.  `QueryAndGroup` lives in `operations.py` (the reference keeps a duplicate in its pointnet2_utils; it is
re-exported here for import compatibility)."""
import torch

from .._ext import sampling
from .operations import QueryAndGroup, ball_query, grouping_operation  # noqa: F401  (re-exports)


class ThreeNN(torch.autograd.Function):
    """(unknown (B, n, 3), known (B, m, 3)) -> (L2 distances to the three nearest known points, ascending
    (B, n, 3); their indices (B, n, 3) int32).  Not differentiable, like the reference's."""

    @staticmethod
    def forward(ctx, unknown, known):
        dist2, idx = sampling.three_nn(unknown.contiguous(), known.contiguous())
        ctx.mark_non_differentiable(idx)
        return dist2.sqrt_(), idx  # the kernel returns squared distances in a buffer of our own

    @staticmethod
    def backward(ctx, *unused):
        return None, None


three_nn = ThreeNN.apply  # type: ignore


class ThreeInterpolate(torch.autograd.Function):
    """(features (B, C, m), idx (B, n, 3) int32, weight (B, n, 3)) -> (B, C, n): for every target point the weighted
    sum of its three source columns.  The gradient flows to `features` only (scatter-add through `idx`)."""

    @staticmethod
    def forward(ctx, features, idx, weight):
        features, idx, weight = features.contiguous(), idx.contiguous(), weight.contiguous()
        batch, channels, m = features.shape
        n = idx.shape[1]
        out = features.new_empty(batch, channels, n)
        sampling.three_interpolate_wrapper(batch, channels, m, n, features, idx, weight, out)
        ctx.save_for_backward(idx, weight)
        ctx.sources = m
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        batch, channels, n = grad_out.shape
        grad_features = grad_out.new_zeros(batch, channels, ctx.sources)
        sampling.three_interpolate_grad_wrapper(batch, channels, n, ctx.sources, grad_out.contiguous(), idx, weight,
                                                grad_features)
        return grad_features, None, None


three_interpolate = ThreeInterpolate.apply  # type: ignore


def propagate_features(unknown, known, known_feats, eps=1e-8):
    """Inverse-distance interpolation of `known_feats` (B, C, m) from the `known` points (B, m, 3) onto the `unknown`
    points (B, n, 3) -> (B, C, n): the three nearest sources of every target, weights 1 / (distance + eps)
    normalised to one -- the interpolation step of `PointnetFPModule.forward`
    (network/pointnet2_modules.py:137-144) as one call."""
    dist, idx = three_nn(unknown, known)
    inv = 1.0 / (dist + eps)
    return three_interpolate(known_feats, idx, inv / inv.sum(dim=2, keepdim=True))


class GroupAll(torch.nn.Module):
    """The whole cloud as ONE group: (xyz (B, N, 3), new_xyz ignored, features (B, C, N) | None) ->
    (B, 3 * use_xyz + C, 1, N) (same result as network/pointnet2_utils.py:127-150)."""

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        parts = []
        if features is None or self.use_xyz:
            parts.append(xyz.transpose(1, 2))
        if features is not None:
            parts.append(features)
        grouped = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
        return grouped.unsqueeze(2)
