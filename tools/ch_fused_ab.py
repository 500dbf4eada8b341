"""Fused (backward inside the resolving kernels) vs separate backward kernels, per shape, default Chamfer path."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import losses
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def med(fn, n=15):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
for B, N in [(32, 2500), (32, 4096), (32, 8192), (64, 8192), (128, 8192), (256, 8192)]:
    a, b = uniform_cloud(B, N, 1).cuda(), uniform_cloud(B, N, 2).cuda()
    d1 = torch.empty(B, N, device="cuda"); d2 = torch.empty(B, N, device="cuda")
    i1 = torch.empty(B, N, dtype=torch.int32, device="cuda"); i2 = torch.empty(B, N, dtype=torch.int32, device="cuda")
    gw = torch.full((2,), 1.0 / (B * N), device="cuda"); g1, g2 = torch.empty_like(a), torch.empty_like(b)
    sums = torch.zeros(2, device="cuda")
    f = med(lambda: losses.nmdistance_forward_backward_uniform(a, b, d1, d2, i1, i2, sums, gw, g1, g2))
    def unf():
        losses.nmdistance_forward(a, b, d1, d2, i1, i2, sums=sums)
        losses.nmdistance_backward_uniform(a, b, g1, g2, gw, i1, i2)
    u = med(unf)
    print("B%d N%d: fused %.4f ms, separate backward %.4f ms" % (B, N, f, u), flush=True)
