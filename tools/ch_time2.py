import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import losses
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(iters):
        flush.zero_()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts)//2]
B, N = int(sys.argv[1]), int(sys.argv[2])
a, b = uniform_cloud(B, N, 1).cuda(), uniform_cloud(B, N, 2).cuda()
bufs = (torch.empty(B, N, device="cuda"), torch.empty(B, N, device="cuda"),
        torch.empty(B, N, dtype=torch.int32, device="cuda"), torch.empty(B, N, dtype=torch.int32, device="cuda"))
for v in [int(x) for x in sys.argv[3].split(",")]:
    for bps in [int(x) for x in sys.argv[4].split(",")]:
        _C.set_option("chamfer_variant", v); _C.set_option("chamfer_blocks_per_sm", bps)
        ms = t(lambda: losses.nmdistance_forward(a, b, *bufs))
        print("B%d N%d variant %d bps %d: fwd (main+finalize, cold L2) %.4f ms" % (B, N, v, bps, ms), flush=True)
