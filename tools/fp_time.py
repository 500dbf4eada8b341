"""three_nn / three_interpolate timing and bit-check against the reference kernels (oracle/_ref) at the FP-stage shape."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
from helpers import uniform_cloud, with_duplicates
from pytorch_points_b200._ext import sampling


def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]


try:
    import ref_sampling
except Exception as e:  # noqa: BLE001
    ref_sampling = None; print("no reference kernels:", e)
for B, N, m in [(16, 16384, 1024), (4, 4097, 1027), (2, 1000, 3), (2, 100, 2), (8, 65536, 4096)]:
    x = with_duplicates(uniform_cloud(B, N, 3)).cuda(); ctr = with_duplicates(uniform_cloud(B, m, 4)).cuda()
    d3 = torch.empty(B, N, 3, device="cuda"); i3 = torch.empty(B, N, 3, dtype=torch.int32, device="cuda")
    ms = timeit(lambda: sampling.three_nn_wrapper(B, N, m, x, ctr, d3, i3))
    line = "three_nn B%d n%d m%d: %.4f ms (%.3e pairs/s)" % (B, N, m, ms, B * N * m / ms * 1e3)
    if ref_sampling is not None:
        rd = torch.empty_like(d3); ri = torch.empty_like(i3)
        rms = timeit(lambda: ref_sampling.three_nn_wrapper(B, N, m, x, ctr, rd, ri), iters=5, warm=1)
        ok = (torch.equal(d3, rd) and torch.equal(i3, ri)) if m >= 3 else (torch.equal(d3[..., :m], rd[..., :m]) and torch.equal(i3[..., :m], ri[..., :m]))
        line += " | reference kernel %.4f ms, bit-equal %s" % (rms, ok)
    print(line, flush=True)
