"""Tensor-core Chamfer sweep (chamfer_variant 51) against the exact kernel (0) and the FFMA sweep (50):
bit equality on a few shapes first (small, so that a hang or a wrong descriptor shows up fast), then times.
    python tools/ch_tc_check.py [quick]"""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud, with_duplicates, lattice_cloud, sphere_cloud
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import losses


def run(a, b, variant):
    B, N, M = a.shape[0], a.shape[1], b.shape[1]
    d1 = torch.full((B, N), -7.0, device="cuda"); d2 = torch.full((B, M), -7.0, device="cuda")
    i1 = torch.full((B, N), -7, dtype=torch.int32, device="cuda"); i2 = torch.full((B, M), -7, dtype=torch.int32, device="cuda")
    sums = torch.zeros(2, device="cuda")
    _C.set_option("chamfer_variant", variant)
    try:
        losses.nmdistance_forward(a, b, d1, d2, i1, i2, sums=sums)
    finally:
        _C.set_option("chamfer_variant", 0)
    torch.cuda.synchronize()
    return d1, d2, i1, i2, sums


bad = 0
cases = [("u", uniform_cloud(1, 128, 1), uniform_cloud(1, 128, 2)), ("u", uniform_cloud(2, 300, 3), uniform_cloud(2, 500, 4)),
         ("u", uniform_cloud(2, 2500, 5), uniform_cloud(2, 2500, 6)), ("u", uniform_cloud(2, 8192, 7), uniform_cloud(2, 8192, 8)),
         ("dups", with_duplicates(uniform_cloud(2, 2500, 9)), with_duplicates(uniform_cloud(2, 3000, 10))),
         ("lattice", lattice_cloud(2, 1500, 11), lattice_cloud(2, 1300, 12)),
         ("offset", uniform_cloud(2, 3000, 13) + 1000.0, uniform_cloud(2, 3000, 14) + 1000.0),
         ("sphere", sphere_cloud(2, 5000, 15), sphere_cloud(2, 5000, 16)),
         ("tiny_scale", uniform_cloud(2, 3000, 17) * 1e-4, uniform_cloud(2, 3000, 18) * 1e-4)]
for name, a, b in cases:
    a, b = a.cuda().contiguous(), b.cuda().contiguous()
    want = run(a, b, 1 if a.shape[1] > 4096 else 32)
    for rep in range(2):
        _C.set_option("chamfer_persistent", rep)
        got = run(a, b, 51)
        _C.set_option("chamfer_persistent", 0)
        nd = [int((x != y).sum()) for x, y in zip(got[:4], want[:4])]
        if any(nd) or not torch.allclose(got[4], want[4], rtol=1e-4):
            bad += 1
            print("MISMATCH %s %s x %s: differing (d1,d2,i1,i2) = %s sums %s vs %s" % (name, tuple(a.shape), tuple(b.shape), nd, got[4].tolist(), want[4].tolist()), flush=True)
            break
    print("case %-10s %s x %s done" % (name, tuple(a.shape), tuple(b.shape)), flush=True)
print("exactness:", "OK" if bad == 0 else "%d problems" % bad, flush=True)
if "quick" in sys.argv:
    sys.exit(1 if bad else 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B, N in [(32, 2500), (32, 4096), (32, 8192), (256, 8192)]:
    a, b = uniform_cloud(B, N, 1).cuda(), uniform_cloud(B, N, 2).cuda()
    d1 = torch.empty(B, N, device="cuda"); d2 = torch.empty(B, N, device="cuda")
    i1 = torch.empty(B, N, dtype=torch.int32, device="cuda"); i2 = torch.empty(B, N, dtype=torch.int32, device="cuda")
    gw = torch.full((2,), 1.0 / (B * N), device="cuda"); g1, g2 = torch.empty_like(a), torch.empty_like(b)
    sums = torch.zeros(2, device="cuda")
    for v, pers in ((1, 0), (51, 0), (51, 1)):
        _C.set_option("chamfer_variant", v)
        _C.set_option("chamfer_persistent", pers)
        fn = lambda: losses.nmdistance_forward_backward_uniform(a, b, d1, d2, i1, i2, sums, gw, g1, g2)
        for _ in range(3): fn()
        torch.cuda.synchronize(); ts = []
        for _ in range(15):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ts.sort(); ms = ts[len(ts) // 2]
        _C.set_option("timing", 1)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        parts = []
        for k in ("chamfer_prep", "chamfer_fwd", "chamfer_finalize"):
            tot, cnt = _C.timing_collect(k)
            if cnt: parts.append("%s %.4f" % (k, tot / cnt))
        _C.set_option("timing", 0)
        print("B%d N%d variant %d persistent %d: fused step %.4f ms (%.3e pairs/s) | %s" % (B, N, v, pers, ms, B * N * N / ms * 1e3, ", ".join(parts)), flush=True)
    _C.set_option("chamfer_variant", 0)
