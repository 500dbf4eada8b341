import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud, sphere_cloud
from pytorch_points_b200 import network as pp
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for maker in (uniform_cloud, sphere_cloud):
    x = maker(16, 16384, 1).cuda()
    ctr = pp.furthest_point_sample(x, 1024, NCHW=False)[1]
    for r, ns in [(0.2, 32), (0.1, 64), (0.05, 32)]:
        print("%s r=%.2f ns=%d: ball_query %.3f ms" % (maker.__name__, r, ns, timeit(lambda: pp.ball_query(r, ns, x, ctr))), flush=True)
