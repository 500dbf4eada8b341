"""Chamfer / nndistance autograd functions.

Drop-in for `pytorch_points.network.model_loss.NmDistanceFunction` / `nndistance`
(network/model_loss.py:401-442) and `LabeledNmdistanceFunction` / `labeled_nndistance`
(:445-483): same inputs, same four outputs `(dist1, dist2, idx1, idx2)`, squared
distances, int32 indices marked non-differentiable.

Differences that do not change results: outputs are allocated directly on the inputs'
device (the reference builds CPU zeros and copies them over, :412-421) and the backward
kernel overwrites its outputs, so no zero fill is needed.
"""
import torch

from .._ext import losses


class NmDistanceFunction(torch.autograd.Function):
    """3D point set to 3D point set distance (both directions in one pass)."""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        xyz1 = xyz1.contiguous()
        xyz2 = xyz2.contiguous()
        assert xyz1.dim() == 3 and xyz2.dim() == 3
        assert xyz1.dtype == xyz2.dtype
        B, n, _ = xyz1.size()
        _, m, _ = xyz2.size()
        dist1 = torch.empty(B, n, dtype=xyz1.dtype, device=xyz1.device)
        dist2 = torch.empty(B, m, dtype=xyz1.dtype, device=xyz1.device)
        idx1 = torch.empty(B, n, dtype=torch.int32, device=xyz1.device)
        idx2 = torch.empty(B, m, dtype=torch.int32, device=xyz1.device)
        losses.nmdistance_forward(xyz1, xyz2, dist1, dist2, idx1, idx2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, graddist1, graddist2, gradNone1, gradNone2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        gradxyz1 = torch.empty_like(xyz1)
        gradxyz2 = torch.empty_like(xyz2)
        losses.nmdistance_backward(xyz1, xyz2, gradxyz1, gradxyz2, graddist1.contiguous(),
                                   graddist2.contiguous(), idx1, idx2)
        return gradxyz1, gradxyz2


nndistance = NmDistanceFunction.apply  # type: ignore


class LabeledNmdistanceFunction(torch.autograd.Function):
    """Chamfer distance restricted to points of the same label; points without a matching
    label get idx -1 and distance 0 and receive no gradient."""

    @staticmethod
    def forward(ctx, xyz1, xyz2, label1, label2):
        xyz1 = xyz1.contiguous()
        xyz2 = xyz2.contiguous()
        assert xyz1.dtype == xyz2.dtype
        B, n, _ = xyz1.size()
        _, m, _ = xyz2.size()
        label1 = label1.to(dtype=xyz1.dtype)
        label2 = label2.to(dtype=xyz1.dtype)
        dist1 = torch.empty(B, n, dtype=xyz1.dtype, device=xyz1.device)
        dist2 = torch.empty(B, m, dtype=xyz1.dtype, device=xyz1.device)
        idx1 = torch.empty(B, n, dtype=torch.int32, device=xyz1.device)
        idx2 = torch.empty(B, m, dtype=torch.int32, device=xyz1.device)
        losses.labeled_nmdistance_forward(xyz1, xyz2, label1, label2, dist1, dist2, idx1, idx2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, graddist1, graddist2, gradNone1, gradNone2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        gradxyz1 = torch.empty_like(xyz1)
        gradxyz2 = torch.empty_like(xyz2)
        losses.nmdistance_backward(xyz1, xyz2, gradxyz1, gradxyz2, graddist1.contiguous(),
                                   graddist2.contiguous(), idx1, idx2)
        return gradxyz1, gradxyz2, None, None


labeled_nndistance = LabeledNmdistanceFunction.apply  # type: ignore


class ChamferSumsFunction(torch.autograd.Function):
    """Fused building block for mean-type Chamfer losses (extension; the reference leaves the
    reduction to user code): returns the 2-vector [sum(dist1), sum(dist2)] produced inside the
    forward kernel's epilogue.  Its backward scatters with two scalar weights instead of two
    (B,N)/(B,M) graddist tensors."""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        xyz1 = xyz1.contiguous()
        xyz2 = xyz2.contiguous()
        B, n, _ = xyz1.size()
        _, m, _ = xyz2.size()
        dev = xyz1.device
        dist1 = torch.empty(B, n, dtype=xyz1.dtype, device=dev)
        dist2 = torch.empty(B, m, dtype=xyz1.dtype, device=dev)
        idx1 = torch.empty(B, n, dtype=torch.int32, device=dev)
        idx2 = torch.empty(B, m, dtype=torch.int32, device=dev)
        sums = torch.empty(2, dtype=xyz1.dtype, device=dev)
        losses.nmdistance_forward(xyz1, xyz2, dist1, dist2, idx1, idx2, sums=sums)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        return sums

    @staticmethod
    def backward(ctx, gsums):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        gradxyz1 = torch.empty_like(xyz1)
        gradxyz2 = torch.empty_like(xyz2)
        losses.nmdistance_backward_uniform(xyz1, xyz2, gradxyz1, gradxyz2, gsums.contiguous(), idx1, idx2)
        return gradxyz1, gradxyz2


chamfer_sums = ChamferSumsFunction.apply  # type: ignore


def chamfer_mean_loss(xyz1, xyz2):
    """mean(dist1) + mean(dist2) with the reduction fused into the kernels (extension)."""
    s = chamfer_sums(xyz1, xyz2)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    return s[0] / (B * n) + s[1] / (B * m)
