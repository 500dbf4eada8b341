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
for (B,N,it) in [(32,8192,5),(4,131072,3),(2,4096,5)]:
    for maker in (uniform_cloud, sphere_cloud):
        p = maker(B,N,4).cuda()
        ref = None
        for prune in (0, 1):
            _C.set_option("knn_prune", prune)
            ms = t(lambda: sampling.knn(16,p,p), it)
            _C.set_option("timing", 1); sampling.knn(16,p,p); torch.cuda.synchronize(); kms, _ = _C.timing_collect("knn"); _C.set_option("timing", 0)
            _C.set_option("knn_stats", 1); out = sampling.knn(16,p,p); v, tot = _C.knn_stats(); _C.set_option("knn_stats", 0)
            if ref is None: ref = out
            else: assert torch.equal(ref[0], out[0]) and torch.equal(ref[1], out[1]), "pruned result differs"
            print("knn k16 B%d N%d %s prune=%d: %.3f ms (sweep kernel %.3f)  %.3g pairs/s  tiles visited %.0f/%.0f = %.1f%%" % (B,N,maker.__name__,prune,ms,kms,B*N*N/ms*1e3,v,tot,100*v/max(tot,1)), flush=True)
_C.set_option("knn_prune", 1)
