"""Sweep kernel alone (serialised launches, library CUDA events), warm and cold L2, one CTA per tile vs persistent."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import losses
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B, N in [(32, 2500), (32, 4096)]:
    a, b = uniform_cloud(B, N, 1).cuda(), uniform_cloud(B, N, 2).cuda()
    d1 = torch.empty(B, N, device="cuda"); d2 = torch.empty(B, N, device="cuda")
    i1 = torch.empty(B, N, dtype=torch.int32, device="cuda"); i2 = torch.empty(B, N, dtype=torch.int32, device="cuda")
    gw = torch.full((2,), 1.0 / (B * N), device="cuda"); g1, g2 = torch.empty_like(a), torch.empty_like(b); sums = torch.zeros(2, device="cuda")
    fn = lambda: losses.nmdistance_forward_backward_uniform(a, b, d1, d2, i1, i2, sums, gw, g1, g2)
    for pers in (0, 1):
        _C.set_option("chamfer_persistent", pers)
        for cold in (0, 1):
            for _ in range(3): fn()
            _C.set_option("timing", 1)
            for nm in ("chamfer_prep", "chamfer_fwd", "chamfer_finalize"): _C.timing_collect(nm)
            for _ in range(10):
                if cold: flush.zero_()
                fn()
            torch.cuda.synchronize()
            parts = {nm: _C.timing_collect(nm) for nm in ("chamfer_prep", "chamfer_fwd", "chamfer_finalize")}
            _C.set_option("timing", 0)
            # whole step with PDL on
            ts = []
            for _ in range(15):
                if cold: flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            ts.sort()
            print("B%d N%d persistent %d %s L2: step %.4f ms | " % (B, N, pers, "cold" if cold else "warm", ts[len(ts) // 2]) +
                  ", ".join("%s %.4f" % (k, v[0] / max(v[1], 1)) for k, v in parts.items()), flush=True)
_C.set_option("chamfer_persistent", -1)
