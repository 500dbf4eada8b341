"""How many points the Chamfer sweep variants (50 / 51) send to the rescan lists."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud, sphere_cloud
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import losses
for maker, nm in ((uniform_cloud, "uniform"), (sphere_cloud, "sphere")):
    for B, N in [(32, 2500), (32, 8192)]:
        a, b = maker(B, N, 1).cuda(), maker(B, N, 2).cuda()
        d1 = torch.empty(B, N, device="cuda"); d2 = torch.empty(B, N, device="cuda")
        i1 = torch.empty(B, N, dtype=torch.int32, device="cuda"); i2 = torch.empty(B, N, dtype=torch.int32, device="cuda")
        for v in (50, 51):
            _C.set_option("chamfer_variant", v)
            losses.nmdistance_forward(a, b, d1, d2, i1, i2)
            torch.cuda.synchronize()
            ws = next(iter(losses._workspaces.values()))
            c = ws[:12 * B].view(torch.int32)
            r2 = c[:B].view(torch.float32)
            print("%s B%d N%d variant %d: ambiguous rows %d cols %d of %d each (%.2f%% / %.2f%%), R^2 max %.3f" % (
                nm, B, N, v, int(c[B:2 * B].sum()), int(c[2 * B:3 * B].sum()), B * N,
                100.0 * int(c[B:2 * B].sum()) / (B * N), 100.0 * int(c[2 * B:3 * B].sum()) / (B * N), float(r2.max())))
        _C.set_option("chamfer_variant", 0)
