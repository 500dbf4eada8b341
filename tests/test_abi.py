"""CPU-only: the C-ABI library loads and exports every function include/pp_b200.h declares,
and the Python binding table matches the header.  No compute calls (no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "pp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pp_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_surface():
    fns = header_functions()
    for must in ["pp_chamfer_fwd", "pp_chamfer_bwd", "pp_chamfer_labeled_fwd", "pp_fps", "pp_gather_fwd",
                 "pp_gather_bwd", "pp_ball_query", "pp_knn", "pp_group_fwd", "pp_group_bwd", "pp_three_nn"]:
        assert must in fns


def test_library_exports_every_declared_symbol():
    from pytorch_points_b200 import _build
    path = _build.build_library()
    lib = ctypes.CDLL(path)
    for fn in header_functions():
        assert hasattr(lib, fn), "libpp_b200.so does not export %s" % fn
    lib.pp_version.restype = ctypes.c_int
    assert lib.pp_version() >= 1


def test_python_binding_table_matches_header():
    from pytorch_points_b200 import _C
    assert sorted(_C.SIGNATURES) == header_functions()


def test_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call, so it can be exercised on CPU."""
    from pytorch_points_b200 import _C
    rc = _C.lib.pp_knn(None, None, 1, 4, 4, 3, 0, None, None, None, 0, 0, None)
    assert rc == -22 and b"k=0" in _C.lib.pp_last_error_string()
    rc = _C.lib.pp_fps(None, 1, 0, 4, 0, None, None, 0, None)
    assert rc == -22
    # at least the packed keys of the exact kernel; the tensor-core sweep's layout (~160 B per point, padded) is larger
    assert _C.lib.pp_chamfer_fwd_workspace_bytes(2, 10, 20) >= 8 * (2 * 10 + 2 * 20)
    need = _C.lib.pp_chamfer_fwd_workspace_bytes(3, 2500, 2500)
    assert 150 * 3 * 5000 <= need <= 180 * 3 * 5120 + 8192
    assert _C.lib.pp_chamfer_fwd_workspace_bytes(0, 10, 10) == 0
    # fused forward + backward: weight vector and gradient pointers are checked up front
    rc = _C.lib.pp_chamfer_fwd_bwd_uniform(None, None, None, 1, 4, 4, None, None, None, None, None, None, None,
                                           None, 0, 0, 0, None)
    assert rc == -22 and b"weight" in _C.lib.pp_last_error_string()
    rc = _C.lib.pp_chamfer_fwd_bwd_uniform(None, None, None, 0, 4, 4, None, None, None, None, None, None, None,
                                           None, 0, 0, 0, None)
    assert rc == 0  # empty batch: nothing to do
    rc = _C.lib.pp_chamfer_labeled_fwd(None, None, None, None, 1, 4, 4, 0, None, None, None, None, None, 0, 0, 0, None)
    assert rc == -22


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from pytorch_points_b200 import network
    with pytest.raises(RuntimeError):
        network.nndistance(torch.rand(1, 4, 3), torch.rand(1, 4, 3))
    with pytest.raises(RuntimeError):
        network.ball_query(0.1, 4, torch.rand(1, 8, 3), torch.rand(1, 2, 3))
    with pytest.raises(RuntimeError):
        network.furthest_point_sample(torch.rand(1, 8, 3), 2, NCHW=False)
    with pytest.raises(RuntimeError):
        network.labeled_nndistance(torch.rand(1, 4, 3), torch.rand(1, 4, 3), torch.zeros(1, 4, 1), torch.zeros(1, 4, 1))
    from pytorch_points_b200._ext import losses
    a = torch.rand(1, 4, 3)
    with pytest.raises(RuntimeError):
        losses.nmdistance_forward_backward_uniform(a, a, torch.empty(1, 4), torch.empty(1, 4),
                                                   torch.empty(1, 4, dtype=torch.int32), torch.empty(1, 4, dtype=torch.int32),
                                                   None, torch.ones(2), torch.empty_like(a), torch.empty_like(a))
