import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import losses
B, N = int(sys.argv[1]), int(sys.argv[2])
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
bps = int(sys.argv[4]) if len(sys.argv) > 4 else 8
fused = len(sys.argv) > 5 and sys.argv[5] == "fused"  # pp_chamfer_fwd_bwd_uniform instead of fwd + bwd
_C.set_option("chamfer_variant", variant); _C.set_option("chamfer_blocks_per_sm", bps)
a, b = uniform_cloud(B, N, 1).cuda(), uniform_cloud(B, N, 2).cuda()
bufs = (torch.empty(B, N, device="cuda"), torch.empty(B, N, device="cuda"),
        torch.empty(B, N, dtype=torch.int32, device="cuda"), torch.empty(B, N, dtype=torch.int32, device="cuda"))
gd1, gd2 = torch.rand(B, N, device="cuda"), torch.rand(B, N, device="cuda")
g1, g2 = torch.empty_like(a), torch.empty_like(b)
gw = torch.full((2,), 1.0 / (B * N), device="cuda")
for _ in range(3):
    if fused:
        losses.nmdistance_forward_backward_uniform(a, b, *bufs, None, gw, g1, g2)
    else:
        losses.nmdistance_forward(a, b, *bufs)
        losses.nmdistance_backward(a, b, g1, g2, gd1, gd2, bufs[2], bufs[3])
torch.cuda.synchronize()
