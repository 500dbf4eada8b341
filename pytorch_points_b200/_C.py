"""ctypes binding of libpp_b200.so (include/pp_b200.h).

The product path has NO fallback: if the CUDA library cannot be loaded, importing this
module raises, and every op raises on non-CUDA tensors.
"""
import ctypes
import os

import torch

from . import _build

_vp = ctypes.c_void_p
_i = ctypes.c_int
_f = ctypes.c_float
_sz = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/pp_b200.h one to one.
SIGNATURES = {
    "pp_version": (_i, []),
    "pp_last_error_string": (ctypes.c_char_p, []),
    "pp_chamfer_fwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "pp_chamfer_last_path": (_i, []),
    "pp_chamfer_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _vp]),
    "pp_chamfer_labeled_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _vp]),
    "pp_chamfer_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "pp_chamfer_bwd_uniform": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "pp_chamfer_fwd_bwd_uniform": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _vp]),
    "pp_fps": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "pp_fps_gather": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "pp_fps_last_plan": (_i, [ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "pp_gather_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _vp]),
    "pp_gather_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _vp]),
    "pp_ball_query": (_i, [_vp, _vp, _i, _i, _i, _f, _i, _vp, _i, _vp]),
    "pp_group_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "pp_group_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "pp_query_group_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _vp, _vp, _i, _vp]),
    "pp_query_group_fwd_pm": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _vp, _vp, _i, _vp]),
    "pp_channels_to_points": (_i, [_vp, _i, _i, _i, _vp, _i, _vp]),
    "pp_query_group_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "pp_knn_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "pp_knn": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _i, _vp]),
    "pp_knn_stats": (_i, [_vp, _vp]),
    "pp_loss_exchange_handle_bytes": (_sz, []),
    "pp_loss_exchange_create": (_i, [_vp, _vp, _i]),
    "pp_loss_exchange_open": (_i, [_vp, _vp, _i]),
    "pp_loss_exchange_bind": (_i, [_vp, _vp, _i, _i]),
    "pp_loss_exchange_close": (_i, [_vp, _vp, _i, _i, _i]),
    "pp_loss_exchange_send": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "pp_loss_exchange_wait": (_i, [_vp, _i, _vp, _vp, _i, _vp]),
    "pp_three_nn": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp]),
    "pp_three_interpolate_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp]),
    "pp_three_interpolate_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp]),
    "pp_microbench": (_i, [_i, _i, ctypes.POINTER(_f), ctypes.POINTER(ctypes.c_double), _i]),
    "pp_set_option": (_i, [ctypes.c_char_p, _i]),
    "pp_dot2": (_i, [_vp, _vp, _vp, _i, _vp]),
    "pp_memcpy_async": (_i, [_vp, _vp, _sz, _i, _vp]),
    "pp_timing_collect": (_i, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_i)]),
}

LIB_PATH = _build.LIB_PATH
PP_CHAMFER_WS_CLEAN = 1


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "pytorch_points_b200: %s is missing. Build it with `python pytorch_points_b200/_build.py` "
            "(or __graft_entry__.build()); there is no CPU or PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header and library disagree
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class PPError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        msg = lib.pp_last_error_string().decode("utf-8", "replace")
        raise PPError("%s failed (code %d): %s" % (what, rc, msg))


def require_cuda(*tensors):
    """The reference asserts CUDA tensors (_ext/utils.h:5-9); so do we -- no CPU path exists."""
    dev = None
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("pytorch_points_b200: expected a CUDA tensor, got device %s "
                               "(there is no CPU implementation of this op)" % t.device)
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("pytorch_points_b200: tensors on different devices (%s vs %s)" % (dev, t.device))
    return dev


def require_contiguous(*tensors):
    for t in tensors:
        if not t.is_contiguous():
            raise RuntimeError("pytorch_points_b200: tensor must be contiguous")


def ptr(t):
    return _vp(t.data_ptr()) if t is not None else _vp(0)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def raw_stream(device):
    """cudaStream_t handle (an int) of torch's current stream on `device`; the raw query is ~10 us cheaper
    per call than building a torch.cuda.Stream object."""
    if _raw_stream is not None:
        return _raw_stream(device.index if device.index is not None else torch.cuda.current_device())
    return torch.cuda.current_stream(device).cuda_stream


def stream_of(device):
    return _vp(raw_stream(device))


def set_option(name, value):
    check(lib.pp_set_option(name.encode(), int(value)), "pp_set_option")


def microbench(which, iters, device=0):
    ms = _f(0)
    work = ctypes.c_double(0)
    check(lib.pp_microbench(int(which), int(iters), ctypes.byref(ms), ctypes.byref(work), int(device)), "pp_microbench")
    return ms.value, work.value


def timing_collect(name):
    """(total_ms, count) of the CUDA-event timings recorded for kernel `name` since the last call
    (requires set_option("timing", 1))."""
    total = ctypes.c_double(0)
    count = _i(0)
    check(lib.pp_timing_collect(name.encode(), ctypes.byref(total), ctypes.byref(count)), "pp_timing_collect")
    return total.value, count.value


def fps_last_plan():
    """(cluster width, points per thread) of the last FPS call made from this thread."""
    c, p = _i(0), _i(0)
    check(lib.pp_fps_last_plan(ctypes.byref(c), ctypes.byref(p)), "pp_fps_last_plan")
    return c.value, p.value


def knn_stats():
    """(tiles_visited, tiles_total) of the last ordered-sweep KNN call made with
    set_option("knn_stats", 1)."""
    v, t = ctypes.c_double(0), ctypes.c_double(0)
    check(lib.pp_knn_stats(ctypes.byref(v), ctypes.byref(t)), "pp_knn_stats")
    return v.value, t.value
