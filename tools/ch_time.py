import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import losses
def t(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(iters):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts)//2]
variants = [int(v) for v in sys.argv[1].split(",")]
sizes = [(32,2500),(32,8192)] if len(sys.argv) < 3 else [tuple(int(v) for v in x.split('x')) for x in sys.argv[2].split(',')]
for (B,N) in sizes:
    a, b = uniform_cloud(B, N, 1).cuda(), uniform_cloud(B, N, 2).cuda()
    bufs = (torch.empty(B, N, device="cuda"), torch.empty(B, N, device="cuda"),
            torch.empty(B, N, dtype=torch.int32, device="cuda"), torch.empty(B, N, dtype=torch.int32, device="cuda"))
    for v in variants:
        _C.set_option("chamfer_variant", v)
        _C.set_option("timing", 1); _C.timing_collect("chamfer_fwd")
        ms = t(lambda: losses.nmdistance_forward(a, b, *bufs))
        tot, cnt = _C.timing_collect("chamfer_fwd"); _C.set_option("timing", 0)
        print("B%d N%d variant %d: fwd call %.4f ms, main kernel %.4f ms (%.3g pairs/s)" % (B, N, v, ms, tot/cnt, B*N*N/(tot/cnt)*1e3), flush=True)
