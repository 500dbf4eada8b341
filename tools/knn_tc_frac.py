"""How many 128 x 128 blocks does the box test keep for nearest-neighbour (k = 1) queries between two independent clouds?"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud, sphere_cloud
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import sampling
_C.set_option("knn_tc", 1)
for B, N, mk in [(32, 8192, uniform_cloud), (32, 8192, sphere_cloud), (32, 2500, uniform_cloud), (32, 4096, uniform_cloud), (8, 16384, uniform_cloud)]:
    a, b = mk(B, N, 1).cuda(), mk(B, N, 2).cuda()
    _C.set_option("knn_stats", 1); sampling.knn(1, a, b); v, tot = _C.knn_stats(); _C.set_option("knn_stats", 0)
    _C.set_option("timing", 1); sampling.knn(1, a, b); torch.cuda.synchronize()
    st = {n: _C.timing_collect(n)[0] for n in ("knn_sort", "knn_prep", "knn_seed", "knn", "knn_select")}
    _C.set_option("timing", 0)
    print("k=1 B%d N%d %s: blocks visited %.1f%%  %s" % (B, N, mk.__name__, 100 * v / tot, " ".join("%s %.3f" % kv for kv in st.items())), flush=True)
