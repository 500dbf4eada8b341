"""Driver for ncu captures of the tensor-core k-NN path: python tools/knn_tc_prof.py [B N k]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import sampling
B, N, k = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (32, 8192, 16)))
p = uniform_cloud(B, N, 4).cuda()
_C.set_option("knn_tc", 1)
for _ in range(3):
    sampling.knn(k, p, p)
torch.cuda.synchronize()
