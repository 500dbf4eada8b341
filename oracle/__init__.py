"""CPU oracle for the pytorch_points hot path -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``oracle/pp_oracle.c`` (a C restatement of the reference's
CUDA kernels, fp32 rounding order spelled out with ``fmaf``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this package; the product (``pytorch_points_b200``)
never does.  Inputs/outputs are numpy arrays (or anything ``np.asarray`` takes).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    """Compile liboracle.so with the Makefile next to this file."""
    src = os.path.join(_HERE, "pp_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-B", "-C", _HERE, "liboracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _f(a):
    a = np.ascontiguousarray(np.asarray(a), dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def _i(a):
    a = np.ascontiguousarray(np.asarray(a), dtype=np.int32)
    return a, a.ctypes.data_as(_i32p)


def chamfer_fwd(xyz1, xyz2):
    """-> dist1 (B,N), dist2 (B,M), idx1 (B,N) int32, idx2 (B,M) int32.  nmdistance_cuda.cu:8-49,118-140"""
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    B, N, c = xyz1.shape
    M = xyz2.shape[1]
    assert xyz2.shape[0] == B and xyz2.shape[2] == c
    d1 = np.empty((B, N), np.float32)
    d2 = np.empty((B, M), np.float32)
    i1 = np.empty((B, N), np.int32)
    i2 = np.empty((B, M), np.int32)
    lib().oracle_chamfer_fwd(p1, p2, B, N, M, c, d1.ctypes.data_as(_f32p), d2.ctypes.data_as(_f32p),
                             i1.ctypes.data_as(_i32p), i2.ctypes.data_as(_i32p))
    return d1, d2, i1, i2


def chamfer_labeled_fwd(xyz1, xyz2, label1, label2):
    """nmdistance_cuda.cu:56-115,142-166.  labels (B,N[,1]) / (B,M[,1]) compared as fp32."""
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    B, N, c = xyz1.shape
    M = xyz2.shape[1]
    label1, l1 = _f(np.asarray(label1).reshape(B, N))
    label2, l2 = _f(np.asarray(label2).reshape(B, M))
    d1 = np.empty((B, N), np.float32)
    d2 = np.empty((B, M), np.float32)
    i1 = np.empty((B, N), np.int32)
    i2 = np.empty((B, M), np.int32)
    lib().oracle_chamfer_labeled_fwd(p1, p2, l1, l2, B, N, M, c, d1.ctypes.data_as(_f32p),
                                     d2.ctypes.data_as(_f32p), i1.ctypes.data_as(_i32p),
                                     i2.ctypes.data_as(_i32p))
    return d1, d2, i1, i2


def chamfer_bwd(xyz1, xyz2, gd1, gd2, idx1, idx2):
    """-> gradxyz1 (B,N,c), gradxyz2 (B,M,c).  nmdistance_cuda.cu:169-221"""
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    gd1, q1 = _f(gd1)
    gd2, q2 = _f(gd2)
    idx1, j1 = _i(idx1)
    idx2, j2 = _i(idx2)
    B, N, c = xyz1.shape
    M = xyz2.shape[1]
    g1 = np.empty((B, N, c), np.float32)
    g2 = np.empty((B, M, c), np.float32)
    lib().oracle_chamfer_bwd(p1, p2, q1, q2, j1, j2, B, N, M, c, g1.ctypes.data_as(_f32p),
                             g2.ctypes.data_as(_f32p))
    return g1, g2


def fps_block_size(n):
    return lib().oracle_fps_block_size(int(n))


def fps(xyz, m, seed=0, return_temp=False):
    """-> idx (B,m) int32 [, temp (B,N)].  sampling_cuda.cu:162-233, geo_operations.py:32-33"""
    xyz, p = _f(xyz)
    B, N, c = xyz.shape
    assert c == 3
    temp = np.full((B, N), 1e10, np.float32)
    idx = np.zeros((B, max(m, 0)), np.int32)
    lib().oracle_fps(p, B, N, int(m), int(seed), temp.ctypes.data_as(_f32p), idx.ctypes.data_as(_i32p))
    return (idx, temp) if return_temp else idx


def ball_query(radius, nsample, xyz, new_xyz):
    """Python-layer argument order of operations.py:90.  -> idx (B,M,nsample) int32.  sampling_cuda.cu:340-376"""
    xyz, p = _f(xyz)
    new_xyz, q = _f(new_xyz)
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = np.empty((B, M, nsample), np.int32)
    lib().oracle_ball_query(q, p, B, N, M, ctypes.c_float(radius), int(nsample), idx.ctypes.data_as(_i32p))
    return idx


def gather_fwd(features, idx):
    """features (B,C,N), idx (B,npoint) -> (B,C,npoint).  sampling_cuda.cu:9-25"""
    features, p = _f(features)
    idx, j = _i(idx)
    B, C, N = features.shape
    npoint = idx.shape[1]
    out = np.empty((B, C, npoint), np.float32)
    lib().oracle_gather_fwd(p, j, B, C, N, npoint, out.ctypes.data_as(_f32p))
    return out


def gather_bwd(grad_out, idx, N):
    """grad_out (B,C,npoint), idx (B,npoint) -> grad_features (B,C,N).  sampling_cuda.cu:47-64"""
    grad_out, p = _f(grad_out)
    idx, j = _i(idx)
    B, C, npoint = grad_out.shape
    g = np.empty((B, C, N), np.float32)
    lib().oracle_gather_bwd(p, j, B, C, N, npoint, g.ctypes.data_as(_f32p))
    return g


def group_fwd(features, idx):
    """features (B,C,N), idx (B,npoint,nsample) -> (B,C,npoint,nsample).  sampling_cuda.cu:447-467"""
    features, p = _f(features)
    idx, j = _i(idx)
    B, C, N = features.shape
    _, npoint, nsample = idx.shape
    out = np.empty((B, C, npoint, nsample), np.float32)
    lib().oracle_group_fwd(p, j, B, C, N, npoint, nsample, out.ctypes.data_as(_f32p))
    return out


def group_bwd(grad_out, idx, N):
    """sampling_cuda.cu:482-503"""
    grad_out, p = _f(grad_out)
    idx, j = _i(idx)
    B, C, npoint, nsample = grad_out.shape
    g = np.empty((B, C, N), np.float32)
    lib().oracle_group_bwd(p, j, B, C, N, npoint, nsample, g.ctypes.data_as(_f32p))
    return g


def knn(k, query, points):
    """query (B,M,c), points (B,N,c) -> dist (B,M,k) ascending, idx (B,M,k) int32; order key (dist, index)."""
    query, q = _f(query)
    points, p = _f(points)
    B, M, c = query.shape
    N = points.shape[1]
    assert k <= N
    dist = np.empty((B, M, k), np.float32)
    idx = np.empty((B, M, k), np.int32)
    lib().oracle_knn(q, p, B, M, N, c, int(k), dist.ctypes.data_as(_f32p), idx.ctypes.data_as(_i32p))
    return dist, idx


def three_nn(unknown, known):
    """unknown (B,N,3), known (B,M,3) -> dist2 (B,N,3) squared, idx (B,N,3).  interpolate_gpu.cu:9-52"""
    unknown, u = _f(unknown)
    known, kn = _f(known)
    B, N, _ = unknown.shape
    M = known.shape[1]
    d = np.empty((B, N, 3), np.float32)
    i = np.empty((B, N, 3), np.int32)
    lib().oracle_three_nn(u, kn, B, N, M, d.ctypes.data_as(_f32p), i.ctypes.data_as(_i32p))
    return d, i


def three_interpolate_fwd(features, idx, weight):
    """features (B,C,M), idx (B,N,3), weight (B,N,3) -> (B,C,N).  interpolate_gpu.cu:77-97"""
    features, p = _f(features)
    idx, j = _i(idx)
    weight, w = _f(weight)
    B, C, M = features.shape
    N = idx.shape[1]
    out = np.empty((B, C, N), np.float32)
    lib().oracle_three_interpolate_fwd(p, j, w, B, C, M, N, out.ctypes.data_as(_f32p))
    return out


def three_interpolate_bwd(grad_out, idx, weight, M):
    """grad_out (B,C,N) -> grad_features (B,C,M).  interpolate_gpu.cu:120-142"""
    grad_out, p = _f(grad_out)
    idx, j = _i(idx)
    weight, w = _f(weight)
    B, C, N = grad_out.shape
    g = np.empty((B, C, M), np.float32)
    lib().oracle_three_interpolate_bwd(p, j, w, B, C, N, M, g.ctypes.data_as(_f32p))
    return g
