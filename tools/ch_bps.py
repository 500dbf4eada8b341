import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import losses
B, N = 32, 2500
a, b = uniform_cloud(B, N, 1).cuda(), uniform_cloud(B, N, 2).cuda()
d1 = torch.empty(B, N, device="cuda"); d2 = torch.empty(B, N, device="cuda")
i1 = torch.empty(B, N, dtype=torch.int32, device="cuda"); i2 = torch.empty(B, N, dtype=torch.int32, device="cuda")
gw = torch.full((2,), 1.0 / (B * N), device="cuda"); g1, g2 = torch.empty_like(a), torch.empty_like(b)
sums = torch.zeros(2, device="cuda"); flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def step():
    losses.nmdistance_forward(a, b, d1, d2, i1, i2, sums=sums)
    losses.nmdistance_backward_uniform(a, b, g1, g2, gw, i1, i2)
def timeit(fn, iters=60):
    for _ in range(5): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
for v in (2, 1):
    for bps in ([int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else (8, 12, 16, 20, 24, 28, 32, 48)):
        _C.set_option("chamfer_variant", v); _C.set_option("chamfer_blocks_per_sm", bps)
        print("variant %d blocks_per_sm %d: step %.4f ms" % (v, bps, timeit(step)), flush=True)
