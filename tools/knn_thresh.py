import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import sampling
def t(fn, iters=7):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(iters):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts)//2]
for (B,N) in [(2,1024),(2,2048),(8,2048),(32,2048),(2,4096),(32,2500),(1,8192),(1,16384)]:
    p = uniform_cloud(B,N,4).cuda()
    res=[]
    for mo in (0,1):
        _C.set_option("knn_morton", mo)
        res.append(t(lambda: sampling.knn(16,p,p)))
    print("B%d N%d k16: streaming %.3f ms | ordered sweep %.3f ms" % (B,N,res[0],res[1]), flush=True)
