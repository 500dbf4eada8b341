"""Tensor-core k-NN path (knn_tc.cu): exactness against the ordered sweep / the oracle, and stage times.
Not part of the product; run on a GPU box:  python tools/knn_tc_check.py [quick]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud, sphere_cloud, with_duplicates, lattice_cloud, np32  # noqa: E402
from pytorch_points_b200 import _C  # noqa: E402
from pytorch_points_b200._ext import sampling  # noqa: E402
import oracle  # noqa: E402

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"


def run(k, q, p, tc):
    _C.set_option("knn_tc", 1 if tc else 0)
    try:
        d, i = sampling.knn(k, q, p)
        torch.cuda.synchronize()
    finally:
        _C.set_option("knn_tc", -1)
    return d, i


def t(fn, iters=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


bad = 0
cases = [
    ("uniform self", 2, 2048, 2048, 16, lambda B, n, s: uniform_cloud(B, n, s)),
    ("uniform M!=N ragged", 3, 1000, 2500, 16, lambda B, n, s: uniform_cloud(B, n, s)),
    ("sphere", 2, 3000, 3000, 8, lambda B, n, s: sphere_cloud(B, n, s)),
    ("duplicates", 2, 2048, 2048, 16, lambda B, n, s: with_duplicates(uniform_cloud(B, n, s))),
    ("lattice", 2, 1500, 1500, 16, lambda B, n, s: lattice_cloud(B, n, s)),
    ("offset 1000", 2, 2048, 2048, 16, lambda B, n, s: uniform_cloud(B, n, s) + 1000.0),
    ("scale 1e-4", 2, 2048, 2048, 16, lambda B, n, s: uniform_cloud(B, n, s) * 1e-4),
    ("tiny N", 2, 300, 40, 32, lambda B, n, s: uniform_cloud(B, n, s)),
    ("k=1", 2, 2048, 2048, 1, lambda B, n, s: uniform_cloud(B, n, s)),
    ("k=32", 1, 4096, 4096, 32, lambda B, n, s: uniform_cloud(B, n, s)),
    ("all equal", 1, 600, 600, 16, lambda B, n, s: torch.zeros(B, n, 3) + 0.25),
    ("clusters", 2, 4096, 4096, 16, lambda B, n, s: (uniform_cloud(B, n, s) * 1e-3 + 100.0 * torch.randint(0, 2, (B, n, 1), generator=torch.Generator().manual_seed(s)).float())),
]
for name, B, M, N, k, mk in cases:
    p = mk(B, N, 11)
    q = p if M == N and "M!=N" not in name else mk(B, M, 12)
    pd = p.cuda()
    qd = pd if q is p else q.cuda()
    d1, i1 = run(k, qd, pd, True)
    ed, ei = oracle.knn(k, np32(q), np32(p))
    okd = np.array_equal(np32(d1), ed)
    oki = np.array_equal(np32(i1), ei)
    print("%-22s B%d M%d N%d k%d: dist %s idx %s" % (name, B, M, N, k, okd, oki), flush=True)
    if not (okd and oki):
        bad += 1
        dd = np32(d1); ii = np32(i1)
        rows = np.argwhere((dd != ed) | (ii != ei))
        print("   first mismatches:", rows[:5].tolist())
        for r in rows[:2]:
            bb, mm, _ = r
            print("   got ", dd[bb, mm, :6], ii[bb, mm, :6]); print("   want", ed[bb, mm, :6], ei[bb, mm, :6])

# non-finite points are never neighbours; a non-finite query gets (inf, -1)
p = uniform_cloud(1, 2048, 5)
p[0, 7] = float("nan"); p[0, 100, 1] = float("inf")
pd = p.cuda()
d1, i1 = run(8, pd, pd, True)
d0, i0 = run(8, pd, pd, False)
same = torch.equal(i1, i0) and torch.equal(d1.isnan(), d0.isnan()) and torch.equal(torch.nan_to_num(d1, 7.0), torch.nan_to_num(d0, 7.0))
print("non-finite points: same as the ordered sweep:", same, "| rows 7:", i1[0, 7, :3].tolist(), d1[0, 7, :3].tolist(), "old:", i0[0, 7, :3].tolist(), d0[0, 7, :3].tolist())
bad += 0 if same else 1

big = [(32, 8192, 16, uniform_cloud), (32, 8192, 16, sphere_cloud), (4, 131072, 16, uniform_cloud), (4, 131072, 16, sphere_cloud), (2, 2048, 16, uniform_cloud), (16, 16384, 16, uniform_cloud)]
if len(sys.argv) > 1 and sys.argv[1] == "regimes":
    big = [(32, 8192, 4, uniform_cloud), (32, 8192, 8, uniform_cloud), (32, 8192, 32, uniform_cloud), (64, 4096, 16, uniform_cloud),
           (128, 2048, 16, uniform_cloud), (8, 32768, 16, uniform_cloud), (1, 262144, 16, uniform_cloud), (32, 2500, 16, sphere_cloud)]
if len(sys.argv) > 1 and sys.argv[1] == "ks":
    big = [(32, 8192, kk, uniform_cloud) for kk in (9, 12, 17, 20, 24, 28)] + [(4, 131072, 17, uniform_cloud), (16, 2048, 17, uniform_cloud)]
if len(sys.argv) > 1 and sys.argv[1] == "small":
    big = [(B_, N_, 16, uniform_cloud) for (B_, N_) in ((1, 2048), (4, 2048), (8, 2048), (16, 2048), (32, 2048), (1, 4096), (4, 4096), (8, 4096), (16, 4096), (1, 8192), (2, 8192), (4, 8192), (8, 8192), (16, 8192), (1, 16384), (4, 16384), (1, 65536))]
if quick:
    big = big[:1]
for B, N, k, mk in big:
    p = mk(B, N, 4).cuda()
    d0, i0 = run(k, p, p, False)
    d1, i1 = run(k, p, p, True)
    ok = torch.equal(d0, d1) and torch.equal(i0, i1)
    bad += 0 if ok else 1
    _C.set_option("knn_tc", 0); ms0 = t(lambda: sampling.knn(k, p, p))
    _C.set_option("knn_tc", 1); ms1 = t(lambda: sampling.knn(k, p, p))
    _C.set_option("timing", 1); sampling.knn(k, p, p); torch.cuda.synchronize()
    st = {n: _C.timing_collect(n)[0] for n in ("knn_sort", "knn_prep", "knn_seed", "knn", "knn_select")}
    _C.set_option("timing", 0)
    _C.set_option("knn_stats", 2); sampling.knn(k, p, p); v, tot = _C.knn_stats(); _C.set_option("knn_stats", 0)
    _C.set_option("knn_tc", -1)
    print("k%d B%d N%d %s: equal=%s  ordered sweep %.3f ms | tensor path %.3f ms  (%s)  blocks visited %.1f%%" % (
        k, B, N, mk.__name__, ok, ms0, ms1, " ".join("%s %.3f" % (n, v_) for n, v_ in st.items()), 100 * v / max(tot, 1)), flush=True)
print("FAILED" if bad else "ALL OK", bad)
