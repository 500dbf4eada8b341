# NOTE: needs libpp_b200.so built with -DPP_FPS_PROFILE (NVCC_FLAGS in _build.py)
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import sampling
x = uniform_cloud(16, 16384, 3).cuda()
idx = torch.empty(16, 1024, dtype=torch.int32, device="cuda")
prof = torch.zeros(8, dtype=torch.int64, device="cuda")
p = prof.data_ptr()
names = ["compute", "warp redux+sts", "bar.sync", "cta redux", "send", "wait", "final pick", "loop top"]
for cl in [1, 2, 4, 8]:
    _C.set_option("fps_cluster", cl)
    temp = torch.full((16, 16384), 1e10, device="cuda")
    sampling.furthest_sampling(1024, 0, x, temp, idx)
    _C.set_option("fps_prof_ptr_lo", (p & 0xffffffff) - (1 << 32) if (p & 0xffffffff) >= (1 << 31) else (p & 0xffffffff))
    _C.set_option("fps_prof_ptr_hi", p >> 32)
    temp = torch.full((16, 16384), 1e10, device="cuda")
    sampling.furthest_sampling(1024, 0, x, temp, idx)
    torch.cuda.synchronize()
    _C.set_option("fps_prof_ptr_lo", 0); _C.set_option("fps_prof_ptr_hi", 0)
    v = prof.cpu().tolist()
    print("cluster", cl, "cycles/round total %.0f" % (sum(v) / 1023), {n: round(c / 1023) for n, c in zip(names, v)})
