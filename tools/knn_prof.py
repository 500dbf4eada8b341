import sys, os, torch
ROOT = "/root/repo"; sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud
from pytorch_points_b200._ext import sampling
B, N = int(sys.argv[1]), int(sys.argv[2])
p = uniform_cloud(B, N, 4).cuda()
for _ in range(3): sampling.knn(16, p, p)
torch.cuda.synchronize()
