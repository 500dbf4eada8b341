import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import sampling
def t(fn, iters=5):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(iters):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts)//2]
for (B,N,it) in [(2,2048,5),(32,8192,5),(4,131072,3)]:
    p = uniform_cloud(B,N,4).cuda()
    for morton in (0, 1):
        for deep in ((0, 1) if morton else (0,)):
          for lex in ((0, 1) if morton else (0,)):
            _C.set_option("knn_morton", morton); _C.set_option("knn_deep_buffers", deep); _C.set_option("knn_lex_only", lex)
            ms = t(lambda: sampling.knn(16,p,p), it)
            print("knn k16 B%d N%d morton=%d deep=%d lexonly=%d: %.3f ms  %.3g pairs/s" % (B,N,morton,deep,lex,ms,B*N*N/ms*1e3), flush=True)
