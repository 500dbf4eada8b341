"""How many points the tensor-core Chamfer path (variant 51) resolves through the block mask (ambiguous:
runner-up granule within TAU of the best value) instead of the single recorded granule."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from helpers import uniform_cloud, sphere_cloud
from _cs_layout import ambiguous, tau
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import losses
for maker, nm in ((uniform_cloud, "uniform"), (sphere_cloud, "sphere")):
    for B, N in [(32, 2500), (32, 8192)]:
        a, b = maker(B, N, 1).cuda(), maker(B, N, 2).cuda()
        d1 = torch.empty(B, N, device="cuda"); d2 = torch.empty(B, N, device="cuda")
        i1 = torch.empty(B, N, dtype=torch.int32, device="cuda"); i2 = torch.empty(B, N, dtype=torch.int32, device="cuda")
        _C.set_option("chamfer_variant", 51)
        losses.nmdistance_forward(a, b, d1, d2, i1, i2)
        _C.set_option("chamfer_variant", 0)
        torch.cuda.synchronize()
        ws = next(iter(losses._workspaces.values()))
        r, c = ambiguous(ws, B, N, N)
        print("%s B%d N%d: ambiguous rows %d cols %d of %d each (%.2f%% / %.2f%%), TAU max %.3e" % (
            nm, B, N, r, c, B * N, 100.0 * r / (B * N), 100.0 * c / (B * N), float(tau(ws, B).max())))
