"""PointNet++ feature-propagation helpers.

Drop-in for `pytorch_points.network.pointnet2_utils.ThreeNN` / `three_nn` and
`ThreeInterpolate` / `three_interpolate` (network/pointnet2_utils.py:11-88) -- SURVEY.md "next"
row N3.  `QueryAndGroup` lives in `operations.py` (the reference keeps a duplicate here,
pointnet2_utils.py:91-124; it is re-exported for import compatibility)."""
import torch

from .._ext import sampling
from .operations import QueryAndGroup, ball_query, grouping_operation  # noqa: F401  (re-exports)


class ThreeNN(torch.autograd.Function):

    @staticmethod
    def forward(ctx, unknown, known):
        """unknown (B, N, 3), known (B, M, 3) -> (dist (B, N, 3) L2 distances to the three nearest
        known points, ascending; idx (B, N, 3) int32)."""
        assert unknown.is_contiguous()
        assert known.is_contiguous()
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = torch.empty(B, N, 3, dtype=torch.float32, device=unknown.device)
        idx = torch.empty(B, N, 3, dtype=torch.int32, device=unknown.device)
        sampling.three_nn_wrapper(B, N, m, unknown, known, dist2, idx)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply  # type: ignore


class ThreeInterpolate(torch.autograd.Function):

    @staticmethod
    def forward(ctx, features, idx, weight):
        """features (B, C, M), idx (B, n, 3), weight (B, n, 3) -> (B, C, n) weighted sum of the three
        gathered feature columns."""
        assert features.is_contiguous()
        assert idx.is_contiguous()
        assert weight.is_contiguous()
        B, c, m = features.size()
        n = idx.size(1)
        ctx.three_interpolate_for_backward = (idx, weight, m)
        output = torch.empty(B, c, n, dtype=torch.float32, device=features.device)
        sampling.three_interpolate_wrapper(B, c, m, n, features, idx, weight, output)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.three_interpolate_for_backward
        B, c, n = grad_out.size()
        grad_features = torch.zeros(B, c, m, dtype=torch.float32, device=grad_out.device)
        sampling.three_interpolate_grad_wrapper(B, c, n, m, grad_out.contiguous(), idx, weight, grad_features)
        return grad_features, None, None


three_interpolate = ThreeInterpolate.apply  # type: ignore


class GroupAll(torch.nn.Module):
    """Groups the whole cloud into one set (network/pointnet2_utils.py:127-150)."""

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        """xyz (B, N, 3), new_xyz ignored, features (B, C, N) -> (B, C + 3, 1, N)."""
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return grouped_xyz
        grouped_features = features.unsqueeze(2)
        if self.use_xyz:
            return torch.cat([grouped_xyz, grouped_features], dim=1)
        return grouped_features
