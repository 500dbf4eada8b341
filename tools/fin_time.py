import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import losses
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (B, N) in [(32, 2500), (32, 8192)]:
    a, b = uniform_cloud(B, N, 1).cuda(), uniform_cloud(B, N, 2).cuda()
    bufs = (torch.empty(B, N, device="cuda"), torch.empty(B, N, device="cuda"),
            torch.empty(B, N, dtype=torch.int32, device="cuda"), torch.empty(B, N, dtype=torch.int32, device="cuda"))
    gw = torch.full((2,), 1e-4, device="cuda"); g1, g2 = torch.empty_like(a), torch.empty_like(b)
    for cold in (0, 1):
        _C.set_option("timing", 1)
        for nm in ("chamfer_fwd", "chamfer_finalize", "chamfer_bwd"): _C.timing_collect(nm)
        for _ in range(20):
            if cold: flush.zero_()
            losses.nmdistance_forward(a, b, *bufs)
            losses.nmdistance_backward_uniform(a, b, g1, g2, gw, bufs[2], bufs[3])
        torch.cuda.synchronize()
        r = {nm: _C.timing_collect(nm) for nm in ("chamfer_fwd", "chamfer_finalize", "chamfer_bwd")}
        _C.set_option("timing", 0)
        print("B%d N%d cold=%d:" % (B, N, cold), {k: round(v[0] / v[1] * 1e3, 1) for k, v in r.items()}, "us", flush=True)
