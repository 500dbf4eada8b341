"""First-contact GPU session: pins the oracle against the reference's own kernels, writes the
golden vectors, measures the roofline denominators MEASURED_PEAKS.json lacks, and times the
kernels (and the reference kernels recompiled for sm_100a) at the headline shapes.
Writes gpurun_out/probe.json.  Not part of the product."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)

from helpers import uniform_cloud  # noqa: E402
from pytorch_points_b200 import _C  # noqa: E402
from pytorch_points_b200._ext import losses, sampling  # noqa: E402

res = {"gpu": torch.cuda.get_device_name(0)}


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


# ---- microbenchmarks
names = ["ffma", "ffma2", "mix_scalar", "mix_packed", "smem", "l2", "redux"]
mb = {}
for w, nm in enumerate(names):
    iters = {0: 4096, 1: 4096, 2: 2048, 3: 2048, 4: 2048, 5: 50, 6: 2048}[w]
    ms, work = _C.microbench(w, iters)
    mb[nm] = {"ms": ms, "work": work, "rate_per_s": work / (ms * 1e-3)}
res["microbench"] = mb
print("microbench", json.dumps(mb, indent=1))


def chamfer_bufs(B, N, M):
    return (torch.empty(B, N, device="cuda"), torch.empty(B, M, device="cuda"),
            torch.empty(B, N, dtype=torch.int32, device="cuda"), torch.empty(B, M, dtype=torch.int32, device="cuda"))


# ---- chamfer timing, all variants
ch = {}
for (B, N) in [(32, 2500), (32, 8192)]:
    a, b = uniform_cloud(B, N, 1).cuda(), uniform_cloud(B, N, 2).cuda()
    bufs = chamfer_bufs(B, N, N)
    for variant in [1, 13, 2, 14]:
        for bps in [24]:
            _C.set_option("chamfer_variant", variant)
            _C.set_option("chamfer_blocks_per_sm", bps)
            med, mn = timeit(lambda: losses.nmdistance_forward(a, b, *bufs))
            ch["B%d_N%d_v%d_bps%d" % (B, N, variant, bps)] = {"ms": med, "min_ms": mn, "pairs_per_s": B * N * N / (med * 1e-3)}
    _C.set_option("chamfer_variant", 0)
    _C.set_option("chamfer_blocks_per_sm", 24)
    gd1, gd2 = torch.rand(B, N, device="cuda"), torch.rand(B, N, device="cuda")
    g1, g2 = torch.empty_like(a), torch.empty_like(b)
    med, mn = timeit(lambda: losses.nmdistance_backward(a, b, g1, g2, gd1, gd2, bufs[2], bufs[3]))
    ch["B%d_N%d_bwd" % (B, N)] = {"ms": med, "min_ms": mn}
res["chamfer"] = ch
print("chamfer", json.dumps(ch, indent=1))

# ---- FPS / ball_query / gather timing (config 3)
x = uniform_cloud(16, 16384, 3).cuda()
fps = {}
for cl in [1, 2, 4, 8]:
    _C.set_option("fps_cluster", cl)
    idx = torch.empty(16, 1024, dtype=torch.int32, device="cuda")

    def run():
        temp = torch.full((16, 16384), 1e10, device="cuda")
        sampling.furthest_sampling(1024, 0, x, temp, idx)
    med, mn = timeit(run, iters=5, warm=2)
    fps["cluster%d" % cl] = {"ms": med, "min_ms": mn, "samples_per_s": 16 * 1024 / (med * 1e-3)}
_C.set_option("fps_cluster", 0)
res["fps"] = fps
print("fps", json.dumps(fps, indent=1))
idx = torch.empty(16, 1024, dtype=torch.int32, device="cuda")
temp = torch.full((16, 16384), 1e10, device="cuda")
sampling.furthest_sampling(1024, 0, x, temp, idx)
ctr = torch.gather(x, 1, idx.long().unsqueeze(-1).expand(16, 1024, 3)).contiguous()
med, mn = timeit(lambda: sampling.ball_query(ctr, x, 0.2, 32))
res["ball_query"] = {"ms": med, "min_ms": mn}
print("ball_query", res["ball_query"])

# ---- KNN timing
kn = {}
for (B, N) in [(2, 2048), (32, 8192)]:
    p = uniform_cloud(B, N, 4).cuda()
    med, mn = timeit(lambda: sampling.knn(16, p, p), iters=5, warm=2)
    kn["B%d_N%d" % (B, N)] = {"ms": med, "min_ms": mn, "pairs_per_s": B * N * N / (med * 1e-3)}
p = uniform_cloud(1, 131072, 5).cuda()
med, mn = timeit(lambda: sampling.knn(16, p, p), iters=2, warm=1)
kn["B1_N131072"] = {"ms": med, "min_ms": mn, "pairs_per_s": 131072.0 ** 2 / (med * 1e-3)}
res["knn"] = kn
print("knn", json.dumps(kn, indent=1))

# ---- the reference's own kernels, recompiled for sm_100a (oracle/_ref)
try:
    import ref_losses as rl
    import ref_sampling as rs
    ref = {}
    for (B, N) in [(32, 2500), (32, 8192)]:
        a, b = uniform_cloud(B, N, 1).cuda(), uniform_cloud(B, N, 2).cuda()
        bufs = chamfer_bufs(B, N, N)
        med, mn = timeit(lambda: rl.nmdistance_forward(a, b, *bufs), iters=5, warm=2)
        ref["chamfer_fwd_B%d_N%d" % (B, N)] = {"ms": med, "pairs_per_s": B * N * N / (med * 1e-3)}
        gd1, gd2 = torch.rand(B, N, device="cuda"), torch.rand(B, N, device="cuda")
        g1, g2 = torch.zeros_like(a), torch.zeros_like(b)
        med, mn = timeit(lambda: rl.nmdistance_backward(a, b, g1, g2, gd1, gd2, bufs[2], bufs[3]), iters=5, warm=2)
        ref["chamfer_bwd_B%d_N%d" % (B, N)] = {"ms": med}
    idx = torch.empty(16, 1024, dtype=torch.int32, device="cuda")

    def run_ref_fps():
        temp = torch.full((16, 16384), 1e10, device="cuda")
        rs.furthest_sampling(1024, 0, x, temp, idx)
    med, mn = timeit(run_ref_fps, iters=3, warm=1)
    ref["fps_16x16384_1024"] = {"ms": med, "samples_per_s": 16 * 1024 / (med * 1e-3)}
    med, mn = timeit(lambda: rs.ball_query(ctr, x, 0.2, 32), iters=5, warm=2)
    ref["ball_query"] = {"ms": med}
    res["reference_cuda"] = ref
    print("reference", json.dumps(ref, indent=1))
except Exception as e:  # noqa: BLE001
    res["reference_cuda"] = {"error": repr(e)}
    print("reference kernels unavailable:", e)

json.dump(res, open(os.path.join(OUT, "probe.json"), "w"), indent=1)
print("wrote", os.path.join(OUT, "probe.json"))
