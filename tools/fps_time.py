import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import sampling
_C.set_option("fps_verbose", 1)
for (B, N, m) in [(16, 16384, 1024), (8, 16384, 1024), (32, 4096, 512), (16, 2048, 512)]:
    x = uniform_cloud(B, N, 3).cuda()
    idx = torch.empty(B, m, dtype=torch.int32, device="cuda")
    for cl in [0, 4, 8]:
        _C.set_option("fps_cluster", cl)
        ts = []
        for i in range(6):
            temp = torch.full((B, N), 1e10, device="cuda")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); sampling.furthest_sampling(m, 0, x, temp, idx); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print("B%d N%d m%d cluster %d: %.3f ms" % (B, N, m, cl, sorted(ts)[len(ts)//2]), flush=True)
