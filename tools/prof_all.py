"""Runs each hot-path kernel a few times (for ncu captures)."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud
from pytorch_points_b200._ext import losses, sampling
which = sys.argv[1]
if which == "fps":
    x = uniform_cloud(16, 16384, 3).cuda()
    idx = torch.empty(16, 1024, dtype=torch.int32, device="cuda")
    for _ in range(2):
        temp = torch.full((16, 16384), 1e10, device="cuda")
        sampling.furthest_sampling(1024, 0, x, temp, idx)
elif which == "bq":
    x = uniform_cloud(16, 16384, 3).cuda()
    idx = torch.empty(16, 1024, dtype=torch.int32, device="cuda")
    temp = torch.full((16, 16384), 1e10, device="cuda")
    sampling.furthest_sampling(1024, 0, x, temp, idx)
    ctr = torch.gather(x, 1, idx.long().unsqueeze(-1).expand(16, 1024, 3)).contiguous()
    for _ in range(2):
        sampling.ball_query(ctr, x, 0.2, 32)
elif which == "knn":
    p = uniform_cloud(int(sys.argv[2]), int(sys.argv[3]), 4).cuda()
    for _ in range(2):
        sampling.knn(16, p, p)
elif which == "qg":
    from pytorch_points_b200 import network as pp
    x = uniform_cloud(16, 16384, 3).cuda()
    f = uniform_cloud(16, 16384, 5, c=64).transpose(1, 2).contiguous().cuda()
    ctr = pp.furthest_point_sample(x, 1024, NCHW=False)[1]
    for _ in range(2):
        pp.query_and_group(x, ctr, f, 0.2, 32, True)
torch.cuda.synchronize()
