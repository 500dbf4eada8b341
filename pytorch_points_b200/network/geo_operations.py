"""Farthest point sampling entry points.

Drop-in for `pytorch_points.network.geo_operations.FurthestPointSampling` and
`furthest_point_sample` (network/geo_operations.py:11-64)."""
import torch

from .._ext import sampling
from .operations import gather_points


class FurthestPointSampling(torch.autograd.Function):

    @staticmethod
    def forward(ctx, xyz, npoint, seedIdx):
        """xyz (B, N, 3) -> idx (B, npoint) int32; idx[:, 0] == seedIdx.  Each next sample is the
        point with the largest distance to the already selected set (reference tie-break)."""
        B, N, _ = xyz.size()
        idx = torch.empty([B, npoint], dtype=torch.int32, device=xyz.device)
        temp = torch.full([B, N], 1e10, dtype=torch.float32, device=xyz.device)
        sampling.furthest_sampling(npoint, seedIdx, xyz, temp, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, grad_idx=None):
        return None, None, None


__furthest_point_sample = FurthestPointSampling.apply  # type: ignore


class FurthestPointSampleGather(torch.autograd.Function):
    """FPS with the gather of the sampled coordinates fused into the sampling kernel."""

    @staticmethod
    def forward(ctx, xyz, npoint, seedIdx):
        """xyz (B, N, 3) contiguous -> (idx (B, npoint) int32, new_xyz (B, npoint, 3))."""
        B, N, _ = xyz.size()
        idx = torch.empty([B, npoint], dtype=torch.int32, device=xyz.device)
        new_xyz = torch.empty([B, npoint, 3], dtype=torch.float32, device=xyz.device)
        temp = torch.full([B, N], 1e10, dtype=torch.float32, device=xyz.device)
        sampling.furthest_sampling_gather(npoint, seedIdx, xyz, temp, idx, new_xyz)
        ctx.save_for_backward(idx)
        ctx.N = N
        ctx.mark_non_differentiable(idx)
        return idx, new_xyz

    @staticmethod
    def backward(ctx, grad_idx, grad_new_xyz):
        # same scatter as GatherFunction.backward on the (B, 3, N) view (network/operations.py:68-85)
        idx, = ctx.saved_tensors
        B, npoint = idx.size()
        g = torch.zeros(B, 3, ctx.N, dtype=torch.float32, device=grad_new_xyz.device)
        sampling.gather_backward(B, 3, ctx.N, npoint, grad_new_xyz.transpose(1, 2).contiguous(), idx, g)
        return g.transpose(1, 2).contiguous(), None, None


__furthest_point_sample_gather = FurthestPointSampleGather.apply  # type: ignore


def furthest_point_sample(xyz, npoint, NCHW=True, seedIdx=0):
    """xyz (B, 3, N) if NCHW else (B, N, 3) -> (idx (B, npoint) int32,
    sampled points (B, 3, npoint) if NCHW else (B, npoint, 3))."""
    assert xyz.dim() == 3, "input for furthest sampling must be a 3D-tensor, but xyz.size() is {}".format(xyz.size())
    if NCHW:
        xyz = xyz.transpose(2, 1).contiguous()
    else:
        xyz = xyz.contiguous()
    assert xyz.size(2) == 3, "furthest sampling is implemented for 3D points"
    # the reference gathers the samples with a second kernel and two transposes
    # (geo_operations.py:60-63); the sampling kernel already holds each winner's coordinates
    idx, sampled_pc = __furthest_point_sample_gather(xyz, npoint, seedIdx)
    if NCHW:
        sampled_pc = sampled_pc.transpose(2, 1).contiguous()
    return idx, sampled_pc


# ---------------------------------------------------------------------------
# k-NN based geometry helpers: the reference's versions call
# pytorch3d.ops.knn_points (network/geo_operations.py:88-152); these run on this repo's KNN.
# ---------------------------------------------------------------------------
def _gather_neighbours(points, idx):
    """points (B, N, C), idx (B, M, K) -> (B, M, K, C)."""
    B, M, K = idx.shape
    C = points.shape[-1]
    flat = torch.gather(points, 1, idx.reshape(B, M * K, 1).long().expand(B, M * K, C))
    return flat.view(B, M, K, C)


def pointUniformLaplacian(points, knn_idx=None, nn_size=3):
    """Uniform (umbrella) Laplacian of a point cloud: point minus the mean of its nn_size nearest
    neighbours.  points (B, N, 3), knn_idx (B, N, K) optional -> (laplacian (B, N, 3), knn_idx)."""
    from .operations import knn_points
    if knn_idx is None:
        _, knn_idx, group = knn_points(points, points, K=nn_size + 1, return_nn=True)
        knn_idx, group = knn_idx[:, :, 1:], group[:, :, 1:, :]
    else:
        group = _gather_neighbours(points, knn_idx)
    lap = points - torch.sum(group, dim=2) / knn_idx.shape[2]
    return lap, knn_idx


def batch_normals(points, base=None, nn_size=20, NCHW=True, idx=None):
    """PCA normals: for every point the direction of least variance of its nn_size nearest
    neighbours in `base` (default: the cloud itself).  points (B, C, M) if NCHW else (B, M, C)
    -> (normals, same layout; idx (B, M, nn_size)).  Sign is arbitrary.  The reference uses its
    cuSOLVER batch_svd extension (out of scope); torch.linalg.svd here."""
    from .operations import knn_points
    if base is None:
        base = points
    if NCHW:
        points = points.transpose(2, 1).contiguous()
        base = base.transpose(2, 1).contiguous()
    assert nn_size < base.shape[1]
    B, M, C = points.shape
    if idx is None:
        _, idx, group = knn_points(points, base, K=nn_size, return_nn=True)
    else:
        group = _gather_neighbours(base, idx)
    centred = group - torch.mean(group, dim=2, keepdim=True)          # (B, M, k, C)
    _, _, vh = torch.linalg.svd(centred.reshape(B * M, nn_size, C), full_matrices=False)
    normals = vh[:, -1, :].reshape(B, M, C)                           # smallest singular direction
    if NCHW:
        normals = normals.transpose(1, 2)
    return normals, idx
