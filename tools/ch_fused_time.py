"""Chamfer step (forward + uniform backward), L2 flushed between steps: four-launch sequence against
pp_chamfer_fwd_bwd_uniform (two launches), eager and as CUDA-graph replays from pinned host buffers."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import losses
from pytorch_points_b200.pipeline import GraphedChamferStep
sizes = [(32, 2500), (32, 8192)] if len(sys.argv) < 2 else [tuple(int(v) for v in x.split('x')) for x in sys.argv[1].split(',')]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, iters=60):
    for _ in range(5): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
for B, N in sizes:
    ah, bh = uniform_cloud(B, N, 1).pin_memory(), uniform_cloud(B, N, 2).pin_memory()
    a, b = ah.cuda(), bh.cuda()
    d1 = torch.empty(B, N, device="cuda"); d2 = torch.empty(B, N, device="cuda")
    i1 = torch.empty(B, N, dtype=torch.int32, device="cuda"); i2 = torch.empty(B, N, dtype=torch.int32, device="cuda")
    gw = torch.full((2,), 1.0 / (B * N), device="cuda"); g1, g2 = torch.empty_like(a), torch.empty_like(b)
    sums = torch.zeros(2, device="cuda")
    def plain():
        losses.nmdistance_forward(a, b, d1, d2, i1, i2, sums=sums)
        losses.nmdistance_backward_uniform(a, b, g1, g2, gw, i1, i2)
    def fused():
        losses.nmdistance_forward_backward_uniform(a, b, d1, d2, i1, i2, sums, gw, g1, g2)
    if True:
        print("B%d N%d eager: four launches %.4f ms, fused %.4f ms" % (B, N, timeit(plain), timeit(fused)), flush=True)
        _C.set_option("timing", 1)
        for nm in ("chamfer_fwd", "chamfer_finalize", "chamfer_bwd"): _C.timing_collect(nm)
        timeit(plain, 20); kp = {nm: _C.timing_collect(nm) for nm in ("chamfer_fwd", "chamfer_finalize", "chamfer_bwd")}
        timeit(fused, 20); kf = {nm: _C.timing_collect(nm) for nm in ("chamfer_fwd", "chamfer_finalize", "chamfer_bwd")}
        _C.set_option("timing", 0)
        fmt = lambda k: ", ".join("%s %.4f" % (n, t / max(c, 1)) for n, (t, c) in k.items())
        print("   kernels (ms): four launches: %s | fused: %s" % (fmt(kp), fmt(kf)), flush=True)
    for f in (False, True):
        st = GraphedChamferStep([(ah, bh)], fused_backward=f)
        for _ in range(10): st.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 200; tk = []
        e0.record()
        for _ in range(K):
            tk.append(st.submit())
            if len(tk) > 1: st.loss(tk.pop(0))
        st.loss(tk.pop(0)); e1.record(); torch.cuda.synchronize()
        print("   graphed e2e (pipelined, from pinned host) fused=%s: %.4f ms/step" % (f, e0.elapsed_time(e1) / K), flush=True)
        del st
