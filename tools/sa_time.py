"""Timing of the set-abstraction sampling+grouping stage: fused kernels vs op-by-op."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud
from pytorch_points_b200 import network as pp


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for (B, N, M, C, r, ns) in [(16, 16384, 1024, 0, 0.2, 32), (16, 16384, 1024, 16, 0.2, 32), (16, 16384, 1024, 64, 0.2, 32),
                            (32, 4096, 512, 64, 0.2, 32), (16, 16384, 1024, 64, 0.1, 64)]:
    xyz = uniform_cloud(B, N, 1).cuda()
    feats = uniform_cloud(B, N, 2, c=C).transpose(1, 2).contiguous().cuda() if C else None
    ctr = pp.furthest_point_sample(xyz, M, NCHW=False)[1]
    fused = pp.QueryAndGroup(r, ns, fused=True)
    comp = pp.QueryAndGroup(r, ns, fused=False)
    assert torch.equal(fused(xyz, ctr, feats), comp(xyz, ctr, feats))
    tf = timeit(lambda: fused(xyz, ctr, feats))
    if C:
        staged = pp.stage_features(feats)
        assert torch.equal(fused(xyz, ctr, feats, staged), comp(xyz, ctr, feats))
        tp = timeit(lambda: fused(xyz, ctr, feats, staged)); tt = timeit(lambda: pp.stage_features(feats))
        print("   staged point-major: grouping %.3f ms (%.0f GB/s out) + staging %.3f ms (%.0f GB/s)" % (
            tp, B * (3 + C) * M * ns * 4 / tp / 1e6, tt, 2 * B * C * N * 4 / tt / 1e6), flush=True)
    tc = timeit(lambda: comp(xyz, ctr, feats))
    tb = timeit(lambda: pp.ball_query(r, ns, xyz, ctr))
    out_bytes = B * (3 + C) * M * ns * 4
    print("B%d N%d M%d C%d r%.2f ns%d: fused %.3f ms (%.0f GB/s out) | op-by-op %.3f ms | ball_query alone %.3f ms"
          % (B, N, M, C, r, ns, tf, out_bytes / tf / 1e6, tc, tb), flush=True)
    ts = timeit(lambda: pp.furthest_point_sample(xyz, M, NCHW=False), n=5)
    print("   fps+gather fused %.3f ms" % ts, flush=True)
