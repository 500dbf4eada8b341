import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud, sphere_cloud
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import sampling
def t(fn, iters=5):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(iters):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts)//2]
for (B,N,it) in [(32,8192,5),(4,131072,3)]:
    for maker in (uniform_cloud, sphere_cloud):
        p = maker(B,N,4).cuda()
        for est in (0, 1):
            _C.set_option("knn_estimate", est)
            ms = t(lambda: sampling.knn(16,p,p), it)
            print("knn k16 B%d N%d %s estimate=%d: %.3f ms  %.3g pairs/s" % (B,N,maker.__name__,est,ms,B*N*N/ms*1e3), flush=True)
