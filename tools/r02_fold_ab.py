"""Round-2 first call: the folded-finalize forward variants (chamfer_variant 41 / 42 / 45, written at
the end of round 1 without GPU time left) against the default path -- exactness first (dist / idx
bit-equal, sums and fused gradients close, key workspace and counters left clean: every shape is
run twice and followed by a default-path call), then the step time, L2 flushed between steps.
    python tools/r02_fold_ab.py            # exits non-zero on any mismatch
"""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud, with_duplicates
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import losses


def run(a, b, variant, fused):
    B, N, M = a.shape[0], a.shape[1], b.shape[1]
    d1 = torch.empty(B, N, device="cuda"); d2 = torch.empty(B, M, device="cuda")
    i1 = torch.empty(B, N, dtype=torch.int32, device="cuda"); i2 = torch.empty(B, M, dtype=torch.int32, device="cuda")
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


bad = 0
for (B, N, M, dup) in [(1, 1, 1, False), (2, 33, 5000, True), (3, 4500, 257, False), (2, 2500, 2500, True),
                       (2, 255, 257, False), (4, 1024, 2048, False), (2, 8192, 8192, False)]:
    a, b = uniform_cloud(B, N, 91).cuda(), uniform_cloud(B, M, 92).cuda()
    if dup:
        a, b = with_duplicates(a.cpu()).cuda(), with_duplicates(b.cpu()).cuda()
    want = run(a, b, 0, False)
    for v in (41, 42, 45):
        for fused in (False, True, True):
            got = run(a, b, v, fused)
            ok = all(torch.equal(x, y) for x, y in zip(got[:4], want[:4]))
            ok = ok and torch.allclose(got[4], want[4], rtol=1e-5)
            for x, y in zip(got[5:], want[5:]):
                ok = ok and float((x - y).abs().max()) <= 1e-5 * float(y.abs().max() + 1e-30)
            if not ok:
                bad += 1
                print("MISMATCH variant %d fused=%s B=%d N=%d M=%d" % (v, fused, B, N, M), flush=True)
        again = run(a, b, 0, False)  # the default path must find keys and counters clean
        if not all(torch.equal(x, y) for x, y in zip(again[:4], want[:4])):
            bad += 1
            print("WORKSPACE LEFT DIRTY by variant %d B=%d N=%d M=%d" % (v, B, N, M), flush=True)
print("exactness:", "OK" if bad == 0 else "%d problems" % bad, flush=True)
if bad:
    sys.exit(1)

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, iters=60):
    for _ in range(5): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
for B, N in [(32, 2500), (32, 2048), (32, 8192)]:
    a, b = uniform_cloud(B, N, 1).cuda(), uniform_cloud(B, N, 2).cuda()
    d1 = torch.empty(B, N, device="cuda"); d2 = torch.empty(B, N, device="cuda")
    i1 = torch.empty(B, N, dtype=torch.int32, device="cuda"); i2 = torch.empty(B, N, dtype=torch.int32, device="cuda")
    gw = torch.full((2,), 1.0 / (B * N), device="cuda"); g1, g2 = torch.empty_like(a), torch.empty_like(b)
    sums = torch.zeros(2, device="cuda")
    for v in (0, 41, 42, 45):
        _C.set_option("chamfer_variant", v)
        ms = timeit(lambda: losses.nmdistance_forward_backward_uniform(a, b, d1, d2, i1, i2, sums, gw, g1, g2))
        print("B%d N%d variant %d: fused step %.4f ms" % (B, N, v, ms), flush=True)
    _C.set_option("chamfer_variant", 0)
