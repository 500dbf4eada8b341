"""Mirror of `pytorch_points._ext.sampling` (_ext/sampling.cpp:205-216)."""
import torch

from .. import _C


def _check(cond, msg):
    if not cond:
        raise RuntimeError("pytorch_points_b200: " + msg)


def furthest_sampling(m, seedIdx, input, temp, idx):
    """sampling.furthest_sampling(m, seedIdx, input, temp, idx) -> idx (_ext/sampling.cpp:68-80)."""
    dev = _C.require_cuda(input, temp, idx)
    _C.require_contiguous(input, temp, idx)
    _check(input.dtype == torch.float32 and temp.dtype == torch.float32, "furthest_sampling: float32 only")
    _check(idx.dtype == torch.int32, "furthest_sampling: idx must be int32")
    B, N, c = input.shape
    _check(c == 3, "furthest sampling is implemented for 3D points")
    rc = _C.lib.pp_fps(_C.ptr(input), B, N, int(m), int(seedIdx), _C.ptr(temp), _C.ptr(idx), dev.index,
                       _C.stream_of(dev))
    _C.check(rc, "pp_fps")
    return idx


def gather_forward(b, c, n, npoints, points, idx, out):
    """sampling.gather_forward (_ext/sampling.cpp:19-28)."""
    dev = _C.require_cuda(points, idx, out)
    _C.require_contiguous(points, idx, out)
    _check(points.dtype == torch.float32 and idx.dtype == torch.int32, "gather_forward: float32 points, int32 idx")
    rc = _C.lib.pp_gather_fwd(_C.ptr(points), _C.ptr(idx), b, c, n, npoints, _C.ptr(out), dev.index,
                              _C.stream_of(dev))
    _C.check(rc, "pp_gather_fwd")
    return 1


def gather_backward(b, c, n, npoints, grad_out, idx, grad_points):
    """sampling.gather_backward (_ext/sampling.cpp:31-41); accumulates into grad_points."""
    dev = _C.require_cuda(grad_out, idx, grad_points)
    _C.require_contiguous(grad_out, idx, grad_points)
    _check(grad_out.dtype == torch.float32 and idx.dtype == torch.int32, "gather_backward: float32 grads, int32 idx")
    rc = _C.lib.pp_gather_bwd(_C.ptr(grad_out), _C.ptr(idx), b, c, n, npoints, _C.ptr(grad_points), dev.index,
                              _C.stream_of(dev))
    _C.check(rc, "pp_gather_bwd")
    return 1


def ball_query(new_xyz, xyz, radius, nsample):
    """sampling.ball_query(new_xyz, xyz, radius, nsample) -> idx (B,M,nsample) int32
    (_ext/sampling.cpp:85-104; the callee allocates the output there too)."""
    dev = _C.require_cuda(new_xyz, xyz)
    _C.require_contiguous(new_xyz, xyz)  # CHECK_INPUT in the reference
    _check(new_xyz.dtype == torch.float32 and xyz.dtype == torch.float32, "ball_query: float32 only")
    B, M, _ = new_xyz.shape
    N = xyz.shape[1]
    idx = torch.empty(B, M, int(nsample), dtype=torch.int32, device=dev)
    rc = _C.lib.pp_ball_query(_C.ptr(new_xyz), _C.ptr(xyz), B, N, M, float(radius), int(nsample), _C.ptr(idx),
                              dev.index, _C.stream_of(dev))
    _C.check(rc, "pp_ball_query")
    return idx


def group_points(points, idx):
    """sampling.group_points(points (B,C,N), idx (B,npoint,nsample)) -> (B,C,npoint,nsample)
    (_ext/sampling.cpp:113-136)."""
    dev = _C.require_cuda(points, idx)
    _C.require_contiguous(points, idx)
    _check(points.dtype == torch.float32, "points must be a float tensor")
    _check(idx.dtype == torch.int32, "idx must be an int tensor")
    B, C, N = points.shape
    _, npoint, nsample = idx.shape
    out = torch.empty(B, C, npoint, nsample, dtype=torch.float32, device=dev)
    rc = _C.lib.pp_group_fwd(_C.ptr(points), _C.ptr(idx), B, C, N, npoint, nsample, _C.ptr(out), dev.index,
                             _C.stream_of(dev))
    _C.check(rc, "pp_group_fwd")
    return out


def group_points_grad(grad_out, idx, n):
    """sampling.group_points_grad(grad_out (B,C,npoint,nsample), idx, n) -> (B,C,n)
    (_ext/sampling.cpp:138-161)."""
    dev = _C.require_cuda(grad_out, idx)
    _C.require_contiguous(grad_out, idx)
    _check(grad_out.dtype == torch.float32, "grad_out must be a float tensor")
    _check(idx.dtype == torch.int32, "idx must be an int tensor")
    B, C, npoint, nsample = grad_out.shape
    g = torch.zeros(B, C, int(n), dtype=torch.float32, device=dev)
    rc = _C.lib.pp_group_bwd(_C.ptr(grad_out), _C.ptr(idx), B, C, int(n), npoint, nsample, _C.ptr(g), dev.index,
                             _C.stream_of(dev))
    _C.check(rc, "pp_group_bwd")
    return g


def three_nn(unknown, known):
    """sampling.three_nn(unknown (B,N,3), known (B,M,3)) -> (dist2 (B,N,3), idx (B,N,3))
    (_ext/sampling.cpp:163-176)."""
    dev = _C.require_cuda(unknown, known)
    _C.require_contiguous(unknown, known)
    _check(unknown.dtype == torch.float32 and known.dtype == torch.float32, "three_nn: float32 only")
    _check(unknown.dim() == 3 and known.dim() == 3 and unknown.shape[2] == 3 and known.shape[2] == 3
           and known.shape[0] == unknown.shape[0], "three_nn: unknown (B,n,3) and known (B,m,3)")
    B, N, _ = unknown.shape
    M = known.shape[1]
    dist2 = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    idx = torch.empty(B, N, 3, dtype=torch.int32, device=dev)
    rc = _C.lib.pp_three_nn(_C.ptr(unknown), _C.ptr(known), B, N, M, _C.ptr(dist2), _C.ptr(idx), dev.index,
                            _C.stream_of(dev))
    _C.check(rc, "pp_three_nn")
    return dist2, idx


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    """sampling.three_nn_wrapper(b, n, m, unknown, known, dist2, idx) with caller-allocated outputs,
    the exact pybind signature of the reference (_ext/sampling.cpp:163-173,213)."""
    dev = _C.require_cuda(unknown, known, dist2, idx)
    _C.require_contiguous(unknown, known, dist2, idx)
    _check(unknown.dtype == torch.float32 and known.dtype == torch.float32 and dist2.dtype == torch.float32,
           "three_nn_wrapper: float32 only")
    _check(idx.dtype == torch.int32, "three_nn_wrapper: idx must be int32")
    _check(unknown.numel() == b * n * 3 and known.numel() == b * m * 3 and dist2.numel() == b * n * 3
           and idx.numel() == b * n * 3, "three_nn_wrapper: tensor sizes disagree with (b, n, m)")
    rc = _C.lib.pp_three_nn(_C.ptr(unknown), _C.ptr(known), b, n, m, _C.ptr(dist2), _C.ptr(idx), dev.index,
                            _C.stream_of(dev))
    _C.check(rc, "pp_three_nn")


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    """sampling.three_interpolate_wrapper(B, c, m, n, points (B,c,m), idx (B,n,3), weight (B,n,3), out (B,c,n))
    (_ext/sampling.cpp:176-188,214)."""
    dev = _C.require_cuda(points, idx, weight, out)
    _C.require_contiguous(points, idx, weight, out)
    _check(idx.dtype == torch.int32, "three_interpolate: idx must be int32")
    _check(points.dtype == torch.float32 and weight.dtype == torch.float32 and out.dtype == torch.float32,
           "three_interpolate: float32 only")
    _check(points.numel() == b * c * m and idx.numel() == b * n * 3 and weight.numel() == b * n * 3
           and out.numel() == b * c * n, "three_interpolate: tensor sizes disagree with (b, c, m, n)")
    rc = _C.lib.pp_three_interpolate_fwd(_C.ptr(points), _C.ptr(idx), _C.ptr(weight), b, c, m, n, _C.ptr(out),
                                         dev.index, _C.stream_of(dev))
    _C.check(rc, "pp_three_interpolate_fwd")


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    """sampling.three_interpolate_grad_wrapper(B, c, n, m, grad_out (B,c,n), idx, weight, grad_points (B,c,m))
    (_ext/sampling.cpp:190-203,215); accumulates into grad_points."""
    dev = _C.require_cuda(grad_out, idx, weight, grad_points)
    _C.require_contiguous(grad_out, idx, weight, grad_points)
    _check(idx.dtype == torch.int32, "three_interpolate_grad: idx must be int32")
    _check(grad_out.dtype == torch.float32 and weight.dtype == torch.float32 and grad_points.dtype == torch.float32,
           "three_interpolate_grad: float32 only")
    _check(grad_out.numel() == b * c * n and idx.numel() == b * n * 3 and weight.numel() == b * n * 3
           and grad_points.numel() == b * c * m, "three_interpolate_grad: tensor sizes disagree with (b, c, n, m)")
    rc = _C.lib.pp_three_interpolate_bwd(_C.ptr(grad_out), _C.ptr(idx), _C.ptr(weight), b, c, n, m,
                                         _C.ptr(grad_points), dev.index, _C.stream_of(dev))
    _C.check(rc, "pp_three_interpolate_bwd")


def knn(k, query, points):
    """group_knn core (no reference counterpart): query (B,M,c), points (B,N,c)
    -> dist (B,M,k) ascending squared distances, idx (B,M,k) int32."""
    dev = _C.require_cuda(query, points)
    _C.require_contiguous(query, points)
    _check(query.dtype == torch.float32 and points.dtype == torch.float32, "knn: float32 only")
    B, M, c = query.shape
    N = points.shape[1]
    _check(points.shape[0] == B and points.shape[2] == c, "knn: query/points shapes disagree")
    dist = torch.empty(B, M, int(k), dtype=torch.float32, device=dev)
    idx = torch.empty(B, M, int(k), dtype=torch.int32, device=dev)
    nbytes = _C.lib.pp_knn_workspace_bytes(B, M, N, c, int(k))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev) if nbytes else None
    rc = _C.lib.pp_knn(_C.ptr(query), _C.ptr(points), B, M, N, c, int(k), _C.ptr(dist), _C.ptr(idx),
                       _C.ptr(ws), nbytes, dev.index, _C.stream_of(dev))
    _C.check(rc, "pp_knn")
    return dist, idx


# ---------------------------------------------------------------------------
# Fused set-abstraction stages (no counterpart in the reference extension: they replace
# Python-level op sequences, see include/pp_b200.h).
# ---------------------------------------------------------------------------
def furthest_sampling_gather(m, seedIdx, input, temp, idx, new_xyz):
    """furthest_sampling + gather of the sampled coordinates into new_xyz (B,m,3) in one launch
    (the op sequence of furthest_point_sample(..., NCHW=False), network/geo_operations.py:44-64)."""
    dev = _C.require_cuda(input, temp, idx, new_xyz)
    _C.require_contiguous(input, temp, idx, new_xyz)
    _check(input.dtype == torch.float32 and temp.dtype == torch.float32 and new_xyz.dtype == torch.float32,
           "furthest_sampling_gather: float32 only")
    _check(idx.dtype == torch.int32, "furthest_sampling_gather: idx must be int32")
    B, N, c = input.shape
    _check(c == 3, "furthest sampling is implemented for 3D points")
    _check(tuple(new_xyz.shape) == (B, int(m), 3), "furthest_sampling_gather: new_xyz must be (B, m, 3)")
    rc = _C.lib.pp_fps_gather(_C.ptr(input), B, N, int(m), int(seedIdx), _C.ptr(temp), _C.ptr(idx),
                              _C.ptr(new_xyz), dev.index, _C.stream_of(dev))
    _C.check(rc, "pp_fps_gather")
    return idx


def channels_to_points(features):
    """(B, C, N) -> (B, N, C) copy: the staging a set-abstraction level does once for all of its scales."""
    dev = _C.require_cuda(features)
    _C.require_contiguous(features)
    _check(features.dtype == torch.float32 and features.dim() == 3, "channels_to_points: float32 (B, C, N)")
    B, C, N = features.shape
    out = torch.empty(B, N, C, dtype=torch.float32, device=dev)
    _C.check(_C.lib.pp_channels_to_points(_C.ptr(features), B, C, N, _C.ptr(out), dev.index, _C.stream_of(dev)),
             "pp_channels_to_points")
    return out


def query_and_group(new_xyz, xyz, features, radius, nsample, use_xyz=True, features_pm=None):
    """ball_query + group xyz (centre-relative) + group features + concat in one kernel
    (QueryAndGroup.forward, network/operations.py:166-213).
    -> (out (B, 3*use_xyz + C, M, nsample), idx (B, M, nsample) int32).
    `features_pm`: the same features staged as (B, N, C) (channels_to_points); read instead of `features`."""
    if features_pm is not None:
        _check(features is not None and features_pm.dim() == 3 and features_pm.dtype == torch.float32 and
               tuple(features_pm.shape) == (features.shape[0], features.shape[2], features.shape[1]),
               "query_and_group: features_pm must be the (B, N, C) copy of features")
        dev = _C.require_cuda(features_pm)
        _C.require_contiguous(features_pm)
    tensors = (new_xyz, xyz) if features is None else (new_xyz, xyz, features)
    dev = _C.require_cuda(*tensors)
    _C.require_contiguous(*tensors)
    _check(all(t.dtype == torch.float32 for t in tensors), "query_and_group: float32 only")
    _check(features is not None or use_xyz, "Cannot have not features and not use xyz as a feature!")
    B, M, _ = new_xyz.shape
    N = xyz.shape[1]
    C = 0 if features is None else features.shape[1]
    if features is not None:
        _check(features.shape[0] == B and features.shape[2] == N, "query_and_group: features must be (B, C, N)")
    ch = (3 if use_xyz else 0) + C
    idx = torch.empty(B, M, int(nsample), dtype=torch.int32, device=dev)
    out = torch.empty(B, ch, M, int(nsample), dtype=torch.float32, device=dev)
    if features_pm is not None and C:
        rc = _C.lib.pp_query_group_fwd_pm(_C.ptr(new_xyz), _C.ptr(xyz), _C.ptr(features_pm), B, N, M, C,
                                          float(radius), int(nsample), int(bool(use_xyz)), _C.ptr(idx), _C.ptr(out),
                                          dev.index, _C.stream_of(dev))
    else:
        rc = _C.lib.pp_query_group_fwd(_C.ptr(new_xyz), _C.ptr(xyz), _C.ptr(features) if C else None, B, N, M, C,
                                       float(radius), int(nsample), int(bool(use_xyz)), _C.ptr(idx), _C.ptr(out),
                                       dev.index, _C.stream_of(dev))
    _C.check(rc, "pp_query_group_fwd")
    return out, idx


def query_and_group_grad(grad_out, idx, n, C, use_xyz, need_features, need_xyz, need_new_xyz):
    """Backward of query_and_group -> (grad_features (B,C,n) | None, grad_xyz (B,n,3) | None,
    grad_new_xyz (B,M,3) | None)."""
    dev = _C.require_cuda(grad_out, idx)
    _C.require_contiguous(grad_out, idx)
    _check(grad_out.dtype == torch.float32 and idx.dtype == torch.int32, "query_and_group_grad: float32 grads, int32 idx")
    B, M, nsample = idx.shape
    gf = torch.zeros(B, C, n, dtype=torch.float32, device=dev) if (need_features and C) else None
    gx = torch.zeros(B, n, 3, dtype=torch.float32, device=dev) if (need_xyz and use_xyz) else None
    gn = torch.empty(B, M, 3, dtype=torch.float32, device=dev) if (need_new_xyz and use_xyz) else None
    rc = _C.lib.pp_query_group_bwd(_C.ptr(grad_out), _C.ptr(idx), B, n, M, C, nsample,
                                   int(bool(use_xyz)), _C.ptr(gf) if gf is not None else None,
                                   _C.ptr(gx) if gx is not None else None, _C.ptr(gn) if gn is not None else None,
                                   dev.index, _C.stream_of(dev))
    _C.check(rc, "pp_query_group_bwd")
    return gf, gx, gn
