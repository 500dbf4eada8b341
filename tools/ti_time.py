"""three_interpolate forward / backward: timing and comparison with the reference kernels (oracle/_ref)."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
from helpers import uniform_cloud
from pytorch_points_b200._ext import sampling
import ref_sampling


def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]


for B, C, m, N in [(16, 64, 1024, 16384), (4, 67, 1027, 4097), (2, 5, 3, 100), (8, 128, 4096, 65536)]:
    g = torch.Generator().manual_seed(1)
    coarse = uniform_cloud(B, m, 6, c=C).transpose(1, 2).contiguous().cuda()
    i3 = torch.randint(0, m, (B, N, 3), generator=g, dtype=torch.int32).cuda()
    w3 = torch.rand(B, N, 3, generator=g).cuda()
    o, ro = torch.empty(B, C, N, device="cuda"), torch.empty(B, C, N, device="cuda")
    ms = timeit(lambda: sampling.three_interpolate_wrapper(B, C, m, N, coarse, i3, w3, o))
    rms = timeit(lambda: ref_sampling.three_interpolate_wrapper(B, C, m, N, coarse, i3, w3, ro), iters=5, warm=1)
    go = torch.rand(B, C, N, generator=g).cuda()
    gp, rgp = torch.zeros(B, C, m, device="cuda"), torch.zeros(B, C, m, device="cuda")
    def ours_b():
        gp.zero_(); sampling.three_interpolate_grad_wrapper(B, C, N, m, go, i3, w3, gp)
    def ref_b():
        rgp.zero_(); ref_sampling.three_interpolate_grad_wrapper(B, C, N, m, go, i3, w3, rgp)
    bms, rbms = timeit(ours_b), timeit(ref_b, iters=5, warm=1)
    print("B%d C%d m%d n%d: fwd %.4f ms (%.0f GB/s out) ref %.4f, bit-equal %s | bwd %.4f ms ref %.4f, max rel diff %.2e" % (
        B, C, m, N, ms, 4.0 * B * C * N / ms / 1e6, rms, torch.equal(o, ro), bms, rbms,
        float(((gp - rgp).abs() / rgp.abs().clamp_min(1e-6)).max())), flush=True)
