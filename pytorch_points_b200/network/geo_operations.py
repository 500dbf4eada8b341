"""Farthest point sampling entry points.

Drop-in for `pytorch_points.network.geo_operations.FurthestPointSampling` and
`furthest_point_sample` (network/geo_operations.py:11-64)."""
import torch

from .._ext import sampling
from .operations import gather_points


def _sample(xyz, npoint, seed, with_points):
    """One farthest-point-sampling call on a contiguous (B, N, 3) cloud: the index buffer and the running minimum
    distances (started at 1e10, like the reference's `temp`) are allocated here; with `with_points` the kernel
    also writes the winners' coordinates (they are in its hands every round)."""
    batch, n = xyz.shape[0], xyz.shape[1]
    idx = xyz.new_empty((batch, npoint), dtype=torch.int32)
    nearest = xyz.new_full((batch, n), 1e10, dtype=torch.float32)
    if not with_points:
        sampling.furthest_sampling(npoint, seed, xyz, nearest, idx)
        return idx, None
    picked = xyz.new_empty((batch, npoint, 3), dtype=torch.float32)
    sampling.furthest_sampling_gather(npoint, seed, xyz, nearest, idx, picked)
    return idx, picked


class FurthestPointSampling(torch.autograd.Function):
    """xyz (B, N, 3) -> idx (B, npoint) int32 with idx[:, 0] == seedIdx; every next sample is the point farthest
    from the set selected so far (the reference's tie-break).  Same signature as the reference's Function
    (network/geo_operations.py:11-40); not differentiable."""

    @staticmethod
    def forward(ctx, xyz, npoint, seedIdx):
        idx, _ = _sample(xyz, npoint, seedIdx, False)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, *unused):
        return None, None, None


class FurthestPointSampleGather(torch.autograd.Function):
    """FPS with the gather of the sampled coordinates fused into the sampling kernel:
    xyz (B, N, 3) contiguous -> (idx (B, npoint) int32, new_xyz (B, npoint, 3)); the gradient of `new_xyz` is
    scattered back onto the sampled points."""

    @staticmethod
    def forward(ctx, xyz, npoint, seedIdx):
        idx, picked = _sample(xyz, npoint, seedIdx, True)
        ctx.save_for_backward(idx)
        ctx.cloud_size = xyz.shape[1]
        ctx.mark_non_differentiable(idx)
        return idx, picked

    @staticmethod
    def backward(ctx, grad_idx, grad_new_xyz):
        # same scatter as GatherFunction.backward on the (B, 3, N) view (network/operations.py:68-85)
        (idx,) = ctx.saved_tensors
        batch, npoint = idx.shape
        grad = grad_new_xyz.new_zeros(batch, 3, ctx.cloud_size)
        sampling.gather_backward(batch, 3, ctx.cloud_size, npoint, grad_new_xyz.transpose(1, 2).contiguous(), idx, grad)
        return grad.transpose(1, 2).contiguous(), None, None


def furthest_point_sample(xyz, npoint, NCHW=True, seedIdx=0):
    """xyz (B, 3, N) if NCHW else (B, N, 3) -> (idx (B, npoint) int32, sampled points in the layout of the input).
    The reference gathers the samples with a second kernel and two transposes (geo_operations.py:60-63); here the
    sampling kernel writes them."""
    if xyz.dim() != 3:
        raise AssertionError("input for furthest sampling must be a 3D-tensor, but xyz.size() is {}".format(xyz.size()))
    cloud = (xyz.transpose(1, 2) if NCHW else xyz).contiguous()
    if cloud.shape[2] != 3:
        raise AssertionError("furthest sampling is implemented for 3D points")
    idx, picked = FurthestPointSampleGather.apply(cloud, npoint, seedIdx)
    return idx, (picked.transpose(1, 2).contiguous() if NCHW else picked)


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
