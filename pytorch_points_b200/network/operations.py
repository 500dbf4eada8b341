"""gather / ball_query / grouping / KNN entry points.

Drop-in for `pytorch_points.network.operations.GatherFunction`, `gather_points`,
`BallQuery`, `ball_query`, `GroupingOperation`, `grouping_operation`, `QueryAndGroup`
(network/operations.py:38-213).  `group_knn` is named by the reference's README (README.md:12)
but absent from the snapshot; its contract is defined here together with a
`knn_points` adaptor shaped like the pytorch3d call the snapshot's callers use
(network/layers.py:52, network/geo_operations.py:112,139)."""
import torch

from .._ext import sampling


class GatherFunction(torch.autograd.Function):
    """features (B, C, N), idx (B, npoint) -> the selected columns (B, C, npoint); the gradient is scattered back
    onto the selected columns.  Same signature as the reference's Function (network/operations.py:38-85)."""

    @staticmethod
    def forward(ctx, features, idx):
        features = features.contiguous()
        idx = idx.to(torch.int32).contiguous()
        (batch, channels, n), npoint = features.shape, idx.shape[1]
        picked = features.new_empty(batch, channels, npoint)
        sampling.gather_forward(batch, channels, n, npoint, features, idx, picked)
        ctx.save_for_backward(idx)
        ctx.source_shape = (batch, channels, n)
        return picked

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        batch, channels, n = ctx.source_shape
        grad = grad_out.new_zeros(batch, channels, n)
        sampling.gather_backward(batch, channels, n, idx.shape[1], grad_out.contiguous(), idx, grad)
        return grad, None


gather_points = GatherFunction.apply  # type: ignore


class BallQuery(torch.autograd.Function):
    """(radius, nsample, xyz (B, N, 3), new_xyz (B, npoint, 3)) -> idx (B, npoint, nsample) int32: the first
    `nsample` points of `xyz` (in index order) within `radius` of each centre, padded with the first hit.
    Argument order of the reference's Function (network/operations.py:88-114); not differentiable."""

    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        members = sampling.ball_query(new_xyz, xyz, radius, nsample)
        ctx.mark_non_differentiable(members)
        return members

    @staticmethod
    def backward(ctx, *unused):
        return (None,) * 4


ball_query = BallQuery.apply  # type: ignore


class GroupingOperation(torch.autograd.Function):
    """features (B, C, N), idx (B, npoint, nsample) -> (B, C, npoint, nsample); the gradient is scatter-added
    through `idx` (network/operations.py:117-163)."""

    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.sources = features.shape[2]
        return sampling.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return sampling.group_points_grad(grad_out.contiguous(), idx, ctx.sources), None


grouping_operation = GroupingOperation.apply  # type: ignore


class QueryAndGroupFunction(torch.autograd.Function):
    """The whole QueryAndGroup stage as one kernel each way (csrc/sa_group.cu)."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, features, radius, nsample, use_xyz, features_pm=None):
        xyz = xyz.contiguous()
        new_xyz = new_xyz.contiguous()
        feats = None if features is None else features.contiguous()
        out, idx = sampling.query_and_group(new_xyz, xyz, feats, radius, nsample, use_xyz, features_pm=features_pm)
        ctx.save_for_backward(idx)
        ctx.meta = (xyz.size(1), 0 if feats is None else feats.size(1), bool(use_xyz))
        ctx.mark_non_differentiable(idx)
        return out, idx

    @staticmethod
    def backward(ctx, grad_out, grad_idx=None):
        idx, = ctx.saved_tensors
        N, C, use_xyz = ctx.meta
        need = ctx.needs_input_grad
        gf, gx, gn = sampling.query_and_group_grad(grad_out.contiguous(), idx, N, C, use_xyz,
                                                   need_features=need[2], need_xyz=need[0], need_new_xyz=need[1])
        return gx, gn, gf, None, None, None, None


def query_and_group(xyz, new_xyz, features, radius, nsample, use_xyz=True, features_pm=None):
    """-> (new_features (B, 3*use_xyz + C, npoint, nsample), idx (B, npoint, nsample) int32).
    `features_pm`: detached (B, N, C) copy of `features` (stage_features) the kernel reads instead; the
    gradient still flows to `features`."""
    return QueryAndGroupFunction.apply(xyz, new_xyz, features, radius, nsample, use_xyz, features_pm)


def stage_features(features):
    """(B, C, N) features -> detached point-major (B, N, C) copy for `QueryAndGroup(..., features_pm=)`:
    done once per set-abstraction level, shared by all of its scales."""
    return sampling.channels_to_points(features.detach().contiguous())


class QueryAndGroup(torch.nn.Module):
    """Ball query around `new_xyz`, then group (centre-relative) coordinates and features.

    Same result as the reference's op sequence (network/operations.py:166-213: ball_query,
    grouping_operation on xyz and on features, centre subtraction, cat), computed by one fused
    kernel; `fused=False` runs the op-by-op sequence instead (kept for parity tests)."""

    def __init__(self, radius, nsample, use_xyz=True, fused=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz, self.fused = radius, nsample, use_xyz, fused

    def forward(self, xyz, new_xyz, features=None, features_pm=None):
        """xyz (B, N, 3), new_xyz (B, npoint, 3), features (B, C, N) -> (B, 3 + C, npoint, nsample).
        `features_pm` (optional, fused path): `stage_features(features)`, computed once by the caller."""
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
        if self.fused:
            return query_and_group(xyz, new_xyz, features, self.radius, self.nsample, self.use_xyz, features_pm)[0]
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            return grouped_xyz
        grouped_features = grouping_operation(features, idx)
        if self.use_xyz:
            return torch.cat([grouped_xyz, grouped_features], dim=1)
        return grouped_features


def _gather_index(idx):
    """int64 copy of a KNN index tensor that is safe to gather / scatter with: a list slot that found
    no neighbour (k > N, or non-finite coordinates -- the kernels only accept d < threshold) carries
    -1 with distance +inf; it is mapped to 0 here instead of tripping a device-side assert."""
    return idx.long().clamp_(min=0)


class _KNNFunction(torch.autograd.Function):
    """dist/idx of the k nearest `points` for every `query` point; distances are differentiable
    w.r.t. both clouds (d/dq = 2 (q - p), d/dp = -2 (q - p))."""

    @staticmethod
    def forward(ctx, k, query, points):
        dist, idx = sampling.knn(k, query, points)
        ctx.save_for_backward(query, points, idx)
        ctx.mark_non_differentiable(idx)
        return dist, idx

    @staticmethod
    def backward(ctx, grad_dist, grad_idx=None):
        query, points, idx = ctx.saved_tensors
        B, M, c = query.shape
        k = idx.shape[2]
        lidx = _gather_index(idx)
        nn = torch.gather(points.unsqueeze(1).expand(B, M, points.shape[1], c), 2,
                          lidx.unsqueeze(-1).expand(B, M, k, c))
        g = 2.0 * grad_dist.unsqueeze(-1) * (query.unsqueeze(2) - nn)  # (B,M,k,c)
        grad_query = g.sum(dim=2)
        grad_points = torch.zeros_like(points)
        grad_points.scatter_add_(1, lidx.reshape(B, M * k, 1).expand(B, M * k, c), -g.reshape(B, M * k, c))
        return None, grad_query, grad_points


def group_knn(k, query, points, unique=True, NCHW=True):
    """k nearest neighbours of `query` among `points`.

    query (B, C, M), points (B, C, N) if NCHW else (B, M, C) / (B, N, C).
    Returns (knn_points, idx, dist): knn_points (B, C, M, k) if NCHW else (B, M, k, C),
    idx (B, M, k) int32, dist (B, M, k) squared L2 ascending; ties broken by lower index.
    `unique` is accepted for signature compatibility with the upstream project's historical
    `group_knn`; duplicates are not removed (pytorch3d.knn_points, which the snapshot's callers
    use, does not remove them either)."""
    if NCHW:
        q = query.transpose(1, 2).contiguous()
        p = points.transpose(1, 2).contiguous()
    else:
        q = query.contiguous()
        p = points.contiguous()
    dist, idx = _KNNFunction.apply(k, q, p)
    B, M, _ = q.shape
    c = p.shape[2]
    nn = torch.gather(p.unsqueeze(1).expand(B, M, p.shape[1], c), 2,
                      _gather_index(idx).unsqueeze(-1).expand(B, M, k, c))  # (B,M,k,C)
    if NCHW:
        nn = nn.permute(0, 3, 1, 2).contiguous()
    return nn, idx, dist


def knn_points(p1, p2, K=1, return_nn=False):
    """Adaptor with the calling convention of `pytorch3d.ops.knn_points` as used by the snapshot
    (e.g. `ops.knn_points(x, x, K=k+1, return_nn=True)`, network/layers.py:52):
    p1 (B, M, C), p2 (B, N, C) -> (dists (B, M, K), idx int64 (B, M, K), nn (B, M, K, C) or None)."""
    dist, idx = _KNNFunction.apply(K, p1.contiguous(), p2.contiguous())
    nn = None
    if return_nn:
        B, M, _ = p1.shape
        c = p2.shape[2]
        nn = torch.gather(p2.unsqueeze(1).expand(B, M, p2.shape[1], c), 2,
                          _gather_index(idx).unsqueeze(-1).expand(B, M, K, c))
    return dist, idx.long(), nn
