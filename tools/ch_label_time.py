"""Labeled Chamfer forward: one-pass kernel with the label mask against the chunk-by-chunk
restatement (chamfer_generic=1) and the unlabeled forward, same clouds.  usage: [BxN,...] [labels]"""
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
sizes = [(32,2500),(32,8192)] if len(sys.argv) < 2 else [tuple(int(v) for v in x.split('x')) for x in sys.argv[1].split(',')]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 4
for (B,N) in sizes:
    a, b = uniform_cloud(B, N, 1).cuda(), uniform_cloud(B, N, 2).cuda()
    g = torch.Generator().manual_seed(3)
    la = torch.randint(0, nl, (B, N), generator=g).float().cuda(); lb = torch.randint(0, nl, (B, N), generator=g).float().cuda()
    bufs = (torch.empty(B, N, device="cuda"), torch.empty(B, N, device="cuda"),
            torch.empty(B, N, dtype=torch.int32, device="cuda"), torch.empty(B, N, dtype=torch.int32, device="cuda"))
    ms_u = t(lambda: losses.nmdistance_forward(a, b, *bufs))
    ms_f = t(lambda: losses.labeled_nmdistance_forward(a, b, la, lb, *bufs))
    _C.set_option("chamfer_generic", 1)
    ms_g = t(lambda: losses.labeled_nmdistance_forward(a, b, la, lb, *bufs), iters=5)
    _C.set_option("chamfer_generic", 0)
    print("B%d N%d labels %d: unlabeled %.4f ms, labeled one-pass %.4f ms (%.3g pairs/s), labeled generic %.4f ms" % (B, N, nl, ms_u, ms_f, B*N*N/ms_f*1e3, ms_g), flush=True)
