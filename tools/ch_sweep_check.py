"""A/B of the Chamfer forward paths: approximate sweep + exact resolution (chamfer_variant 0) against
the exact one-pass kernel (chamfer_variant 60) -- bit equality of dist / idx on plain, tied, lattice,
offset and clustered inputs, fused gradients, then step times with the L2 flushed.
    python tools/ch_sweep_check.py [quick]
"""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud, with_duplicates, lattice_cloud, sphere_cloud
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import losses


def run(a, b, variant, fused):
    B, N, M = a.shape[0], a.shape[1], b.shape[1]
    d1 = torch.full((B, N), -7.0, device="cuda"); d2 = torch.full((B, M), -7.0, device="cuda")
    i1 = torch.full((B, N), -7, dtype=torch.int32, device="cuda"); i2 = torch.full((B, M), -7, dtype=torch.int32, device="cuda")
    sums = torch.zeros(2, device="cuda"); gw = torch.tensor([0.5 / (B * N), 2.0 / (B * M)], device="cuda")
    g1, g2 = torch.full_like(a, 3.0), torch.full_like(b, -3.0)
    _C.set_option("chamfer_variant", variant)
    try:
        if fused:
            losses.nmdistance_forward_backward_uniform(a, b, d1, d2, i1, i2, sums, gw, g1, g2)
        else:
            losses.nmdistance_forward(a, b, d1, d2, i1, i2, sums=sums)
            losses.nmdistance_backward_uniform(a, b, g1, g2, gw, i1, i2)
    finally:
        _C.set_option("chamfer_variant", 0)
    torch.cuda.synchronize()
    return d1, d2, i1, i2, sums, g1, g2


def clustered(B, N, seed):
    g = torch.Generator().manual_seed(seed)
    centres = torch.rand(B, 8, 3, generator=g) * 100.0
    pick = torch.randint(0, 8, (B, N), generator=g)
    return torch.gather(centres, 1, pick.unsqueeze(-1).expand(B, N, 3)) + 1e-3 * torch.rand(B, N, 3, generator=g)


cases = []
for (B, N, M) in [(1, 1, 1), (2, 33, 5000), (3, 4500, 257), (2, 2500, 2500), (2, 255, 257), (4, 1024, 2048),
                  (1, 129, 127), (2, 8192, 8192), (3, 700, 300)]:
    cases.append(("uniform", uniform_cloud(B, N, 91), uniform_cloud(B, M, 92)))
cases.append(("dups", with_duplicates(uniform_cloud(2, 2500, 5)), with_duplicates(uniform_cloud(2, 3000, 6))))
cases.append(("lattice", lattice_cloud(2, 1500, 7), lattice_cloud(2, 1300, 8)))
cases.append(("same", uniform_cloud(2, 2000, 9), uniform_cloud(2, 2000, 9)))
cases.append(("offset1e3", uniform_cloud(2, 3000, 10) + 1000.0, uniform_cloud(2, 3000, 11) + 1000.0))
cases.append(("offset_far", uniform_cloud(2, 3000, 12) + 50.0, uniform_cloud(2, 3000, 13) - 50.0))
cases.append(("clustered", clustered(2, 4000, 14), clustered(2, 4000, 14) + 1e-4))
cases.append(("tiny_scale", uniform_cloud(2, 3000, 15) * 1e-4, uniform_cloud(2, 3000, 16) * 1e-4))
cases.append(("sphere", sphere_cloud(2, 5000, 17), sphere_cloud(2, 5000, 18)))
cases.append(("all_equal", torch.ones(2, 600, 3), torch.ones(2, 500, 3)))
bad = 0
for name, a, b in cases:
    a, b = a.cuda().contiguous(), b.cuda().contiguous()
    want = run(a, b, 0, False)
    for fused in (False, True, True):
        got = run(a, b, 50, fused)
        ok = all(torch.equal(x, y) for x, y in zip(got[:4], want[:4]))
        ok = ok and torch.allclose(got[4], want[4], rtol=1e-4)
        for x, y in zip(got[5:], want[5:]):
            ok = ok and float((x - y).abs().max()) <= 1e-5 * float(y.abs().max() + 1e-30)
        if not ok:
            bad += 1
            nd = [int((x != y).sum()) for x, y in zip(got[:4], want[:4])]
            print("MISMATCH %s fused=%s shape %s %s: differing (d1,d2,i1,i2) = %s sums %s vs %s" % (
                name, fused, tuple(a.shape), tuple(b.shape), nd, got[4].tolist(), want[4].tolist()), flush=True)
    print("case %-10s %s x %s done" % (name, tuple(a.shape), tuple(b.shape)), flush=True)
print("exactness:", "OK" if bad == 0 else "%d problems" % bad, flush=True)
if bad or "quick" in sys.argv:
    sys.exit(1 if bad else 0)

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, iters=40):
    for _ in range(5): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
for B, N in [(32, 2500), (32, 2048), (32, 8192), (256, 8192)]:
    a, b = uniform_cloud(B, N, 1).cuda(), uniform_cloud(B, N, 2).cuda()
    d1 = torch.empty(B, N, device="cuda"); d2 = torch.empty(B, N, device="cuda")
    i1 = torch.empty(B, N, dtype=torch.int32, device="cuda"); i2 = torch.empty(B, N, dtype=torch.int32, device="cuda")
    gw = torch.full((2,), 1.0 / (B * N), device="cuda"); g1, g2 = torch.empty_like(a), torch.empty_like(b)
    sums = torch.zeros(2, device="cuda")
    for v in (0, 50):
        _C.set_option("chamfer_variant", v)
        ms = timeit(lambda: losses.nmdistance_forward_backward_uniform(a, b, d1, d2, i1, i2, sums, gw, g1, g2))
        _C.set_option("timing", 1)
        for _ in range(5):
            losses.nmdistance_forward_backward_uniform(a, b, d1, d2, i1, i2, sums, gw, g1, g2)
        torch.cuda.synchronize()
        parts = []
        for k in ("chamfer_prep", "chamfer_fwd", "chamfer_finalize", "chamfer_rescan"):
            tot, cnt = _C.timing_collect(k)
            if cnt: parts.append("%s %.4f" % (k, tot / cnt))
        _C.set_option("timing", 0)
        print("B%d N%d variant %d: fused step %.4f ms (%.3e pairs/s) | %s" % (B, N, v, ms, B * N * N / ms * 1e3, ", ".join(parts)), flush=True)
    _C.set_option("chamfer_variant", 0)
