"""Mirror of `pytorch_points._ext.losses` (_ext/nmdistance.cpp:30-34).

The library switches to the tensors' device itself (DeviceGuard in csrc), so no Python-side
device context is needed around the calls.

Caller allocates every output, exactly like the reference; functions return 1 on success
(the reference returns 1 / 0 and prints on failure -- here failures raise instead)."""
import torch

from .. import _C

_workspaces = {}   # (device index, stream handle) -> current scratch tensor
_retired = []      # outgrown scratch tensors: kept alive, a CUDA graph may have baked their address in


def _workspace(device, nbytes, stream_handle=None):
    """Scratch for the Chamfer forward (prepared clouds, packed keys, candidate lists).  The library
    initialises what it reads, so the buffer carries no state between calls (flags = 0).
    One buffer per (device, stream): calls on one stream never overlap.  A buffer that has been
    handed out is never freed -- when a larger shape arrives the old tensor moves to `_retired`, so a
    CUDA graph captured around an earlier call keeps a valid address.  While the current stream is
    being captured the call gets a fresh tensor of its own from the graph's private pool: replays of
    different graphs (or a replay racing an eager call) then never share scratch."""
    if torch.cuda.is_current_stream_capturing():
        return torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=device)
    if stream_handle is None:
        stream_handle = _C.raw_stream(device)
    key = (device.index, stream_handle)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _retired.append(buf)
        grow = nbytes if buf is None else max(nbytes, buf.numel() * 3 // 2)
        buf = torch.empty((max(grow, 1 << 20),), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def _check_f32(*ts):
    for t in ts:
        if t.dtype != torch.float32:
            raise RuntimeError("pytorch_points_b200: only float32 point clouds are supported, got %s" % t.dtype)


_ws_bytes = {}


def _scratch(dev, B, N, M, workspace, workspace_clean, stream_handle=None):
    """(tensor, flags) for a Chamfer forward: the caller's own buffer (`workspace`, a CUDA uint8 tensor of
    at least pp_chamfer_fwd_workspace_bytes; `workspace_clean` = the caller filled it with 0xff once and
    only ever passes it to these functions on one stream -> PP_CHAMFER_WS_CLEAN, no per-call fill), or the
    module's per-stream buffer with flags 0."""
    nbytes = _ws_bytes.get((B, N, M))
    if nbytes is None:
        if len(_ws_bytes) > 256:
            _ws_bytes.clear()
        nbytes = _ws_bytes[(B, N, M)] = _C.lib.pp_chamfer_fwd_workspace_bytes(B, N, M)
    if workspace is None:
        return _workspace(dev, nbytes, stream_handle), 0
    if not workspace.is_cuda or workspace.device != dev or workspace.dtype != torch.uint8 \
            or not workspace.is_contiguous() or workspace.numel() < nbytes:
        raise RuntimeError("chamfer workspace must be a contiguous CUDA uint8 tensor of >= %d bytes on %s" % (nbytes, dev))
    return workspace, (_C.PP_CHAMFER_WS_CLEAN if workspace_clean else 0)


def nmdistance_forward(xyz1, xyz2, dist1, dist2, idx1, idx2, sums=None, workspace=None, workspace_clean=False):
    """losses.nmdistance_forward(xyz1, xyz2, dist1, dist2, idx1, idx2) -> int
    (_ext/nmdistance.cpp:13-15).  `sums` (optional, 2 floats) is an extension: fused
    [sum(dist1), sum(dist2)]; `workspace` / `workspace_clean`: see _scratch."""
    dev = _C.require_cuda(xyz1, xyz2, dist1, dist2, idx1, idx2)
    _C.require_contiguous(xyz1, xyz2, dist1, dist2, idx1, idx2)
    _check_f32(xyz1, xyz2, dist1, dist2)
    B, N, c = xyz1.shape
    M = xyz2.shape[1]
    if xyz2.shape[0] != B or xyz2.shape[2] != c:
        raise RuntimeError("nmdistance_forward: xyz1 %s and xyz2 %s disagree" % (tuple(xyz1.shape), tuple(xyz2.shape)))
    if idx1.dtype != torch.int32 or idx2.dtype != torch.int32:
        raise RuntimeError("nmdistance_forward: idx tensors must be int32")
    if dist1.numel() != B * N or idx1.numel() != B * N or dist2.numel() != B * M or idx2.numel() != B * M:
        raise RuntimeError("nmdistance_forward: outputs must be (B,N) / (B,M)")
    if sums is not None:
        _C.require_cuda(xyz1, sums)
        if sums.dtype != torch.float32 or sums.numel() != 2 or not sums.is_contiguous():
            raise RuntimeError("nmdistance_forward: sums must be 2 contiguous float32 values")
    stream = _C.raw_stream(dev)
    ws, flags = _scratch(dev, B, N, M, workspace, workspace_clean, stream)
    rc = _C.lib.pp_chamfer_fwd(_C.ptr(xyz1), _C.ptr(xyz2), B, N, M, c, _C.ptr(dist1), _C.ptr(dist2),
                               _C.ptr(idx1), _C.ptr(idx2), _C.ptr(sums), _C.ptr(ws), ws.numel(),
                               flags, dev.index, stream)
    _C.check(rc, "pp_chamfer_fwd")
    return 1


def labeled_nmdistance_forward(xyz1, xyz2, label1, label2, dist1, dist2, idx1, idx2):
    """losses.labeled_nmdistance_forward (_ext/nmdistance.cpp:17-20)."""
    dev = _C.require_cuda(xyz1, xyz2, label1, label2, dist1, dist2, idx1, idx2)
    label1 = label1.to(dtype=xyz1.dtype).contiguous()  # nmdistance_cuda.cu:153 (toType)
    label2 = label2.to(dtype=xyz1.dtype).contiguous()
    _C.require_contiguous(xyz1, xyz2, dist1, dist2, idx1, idx2)
    _check_f32(xyz1, xyz2, dist1, dist2)
    B, N, c = xyz1.shape
    M = xyz2.shape[1]
    if label1.numel() != B * N or label2.numel() != B * M:
        raise RuntimeError("labeled_nmdistance_forward: labels must be (B,N[,1]) and (B,M[,1])")
    if idx1.dtype != torch.int32 or idx2.dtype != torch.int32:
        raise RuntimeError("labeled_nmdistance_forward: idx tensors must be int32")
    if dist1.numel() != B * N or idx1.numel() != B * N or dist2.numel() != B * M or idx2.numel() != B * M:
        raise RuntimeError("labeled_nmdistance_forward: outputs must be (B,N) / (B,M)")
    nbytes = _C.lib.pp_chamfer_fwd_workspace_bytes(B, N, M)
    ws = _workspace(dev, nbytes)
    rc = _C.lib.pp_chamfer_labeled_fwd(_C.ptr(xyz1), _C.ptr(xyz2), _C.ptr(label1), _C.ptr(label2), B, N, M, c,
                                       _C.ptr(dist1), _C.ptr(dist2), _C.ptr(idx1), _C.ptr(idx2),
                                       _C.ptr(ws), ws.numel(), 0, dev.index, _C.stream_of(dev))
    _C.check(rc, "pp_chamfer_labeled_fwd")
    return 1


def nmdistance_backward(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2):
    """losses.nmdistance_backward (_ext/nmdistance.cpp:23-27).  gradxyz1/2 are overwritten."""
    dev = _C.require_cuda(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2)
    _C.require_contiguous(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2)
    _check_f32(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2)
    B, N, c = xyz1.shape
    M = xyz2.shape[1]
    if idx1.dtype != torch.int32 or idx2.dtype != torch.int32:
        raise RuntimeError("nmdistance_backward: idx tensors must be int32")
    if gradxyz1.shape != xyz1.shape or gradxyz2.shape != xyz2.shape or graddist1.numel() != B * N \
            or graddist2.numel() != B * M or idx1.numel() != B * N or idx2.numel() != B * M:
        raise RuntimeError("nmdistance_backward: shapes disagree")
    rc = _C.lib.pp_chamfer_bwd(_C.ptr(xyz1), _C.ptr(xyz2), _C.ptr(graddist1), _C.ptr(graddist2),
                               _C.ptr(idx1), _C.ptr(idx2), B, N, M, c, _C.ptr(gradxyz1), _C.ptr(gradxyz2),
                               dev.index, _C.stream_of(dev))
    _C.check(rc, "pp_chamfer_bwd")
    return 1


def nmdistance_backward_uniform(xyz1, xyz2, gradxyz1, gradxyz2, gw, idx1, idx2):
    """Extension: backward for a loss that depends on dist1/dist2 only through
    [sum(dist1), sum(dist2)]; `gw` is the 2-element CUDA tensor of upstream gradients of those
    sums.  No (B,N)/(B,M) graddist tensors are materialised."""
    dev = _C.require_cuda(xyz1, xyz2, gradxyz1, gradxyz2, gw, idx1, idx2)
    _C.require_contiguous(xyz1, xyz2, gradxyz1, gradxyz2, gw, idx1, idx2)
    _check_f32(xyz1, xyz2, gradxyz1, gradxyz2, gw)
    if gw.numel() != 2:
        raise RuntimeError("nmdistance_backward_uniform: gw must hold 2 floats")
    B, N, c = xyz1.shape
    M = xyz2.shape[1]
    rc = _C.lib.pp_chamfer_bwd_uniform(_C.ptr(xyz1), _C.ptr(xyz2), _C.ptr(gw), _C.ptr(idx1), _C.ptr(idx2),
                                       B, N, M, c, _C.ptr(gradxyz1), _C.ptr(gradxyz2), dev.index,
                                       _C.stream_of(dev))
    _C.check(rc, "pp_chamfer_bwd_uniform")
    return 1


def nmdistance_forward_backward_uniform(xyz1, xyz2, dist1, dist2, idx1, idx2, sums, gw, gradxyz1, gradxyz2,
                                        workspace=None, workspace_clean=False):
    """Extension: `nmdistance_forward(..., sums=sums)` followed by `nmdistance_backward_uniform` in
    two launches instead of four -- the backward rides in the kernel that resolves the indices.
    For losses that depend on dist1/dist2 only through their sums with weights `gw` known up
    front (mean / sum Chamfer).  `sums` may be None.  c == 3."""
    dev = _C.require_cuda(xyz1, xyz2, dist1, dist2, idx1, idx2, gw, gradxyz1, gradxyz2)
    _C.require_contiguous(xyz1, xyz2, dist1, dist2, idx1, idx2, gw, gradxyz1, gradxyz2)
    _check_f32(xyz1, xyz2, dist1, dist2, gw, gradxyz1, gradxyz2)
    B, N, c = xyz1.shape
    M = xyz2.shape[1]
    if c != 3 or xyz2.shape[0] != B or xyz2.shape[2] != 3:
        raise RuntimeError("nmdistance_forward_backward_uniform: needs (B,N,3) and (B,M,3) clouds")
    if idx1.dtype != torch.int32 or idx2.dtype != torch.int32:
        raise RuntimeError("nmdistance_forward_backward_uniform: idx tensors must be int32")
    if gw.numel() != 2 or gradxyz1.shape != xyz1.shape or gradxyz2.shape != xyz2.shape:
        raise RuntimeError("nmdistance_forward_backward_uniform: gw must hold 2 floats, gradients match the clouds")
    stream = _C.raw_stream(dev)
    ws, flags = _scratch(dev, B, N, M, workspace, workspace_clean, stream)
    rc = _C.lib.pp_chamfer_fwd_bwd_uniform(_C.ptr(xyz1), _C.ptr(xyz2), _C.ptr(gw), B, N, M, _C.ptr(dist1),
                                           _C.ptr(dist2), _C.ptr(idx1), _C.ptr(idx2), _C.ptr(sums),
                                           _C.ptr(gradxyz1), _C.ptr(gradxyz2), _C.ptr(ws), ws.numel(),
                                           flags, dev.index, stream)
    _C.check(rc, "pp_chamfer_fwd_bwd_uniform")
    return 1


def _fwd_bwd_uniform_raw(dev, B, N, M, xyz1, xyz2, fbase, ibase, gw, g1, g2):
    """Host-lean form of nmdistance_forward_backward_uniform for callers that allocated every output
    themselves (network.model_loss): `fbase` / `ibase` are the addresses of one float buffer laid out
    dist1 (B*N) | dist2 (B*M) | sums (2) and one int32 buffer idx1 | idx2.  No per-call validation, no
    tensor views: ~20 us less host time per step, which is what a 0.1 ms GPU step is bounded by."""
    stream = _C.raw_stream(dev)
    ws, flags = _scratch(dev, B, N, M, None, False, stream)
    vp = _C._vp
    rc = _C.lib.pp_chamfer_fwd_bwd_uniform(vp(xyz1.data_ptr()), vp(xyz2.data_ptr()), vp(gw.data_ptr()), B, N, M,
                                           vp(fbase), vp(fbase + 4 * B * N), vp(ibase), vp(ibase + 4 * B * N),
                                           vp(fbase + 4 * B * (N + M)), vp(g1.data_ptr()), vp(g2.data_ptr()),
                                           vp(ws.data_ptr()), ws.numel(), flags, dev.index, stream)
    _C.check(rc, "pp_chamfer_fwd_bwd_uniform")
