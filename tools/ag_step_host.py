"""Host microseconds per piece of the autograd end-to-end step at B=32, N=M=2500 (GPU step ~0.09 ms: host bound)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud
from pytorch_points_b200.dist import sharded_chamfer_loss
from pytorch_points_b200.pipeline import HostPrefetcher, HostScalarReader
B, N = 32, 2500
dev = torch.device("cuda")
ah, bh = uniform_cloud(B, N, 1).pin_memory(), uniform_cloud(B, N, 2).pin_memory()
pf, rd = HostPrefetcher(dev, depth=2), HostScalarReader(dev, depth=4)
pf.prefetch((ah, bh))
acc = {}
def tick(name, t0):
    t1 = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t1 - t0); return t1
def step(measure):
    t = time.perf_counter()
    xd, yd = pf.get();                                              t = tick("prefetcher.get", t) if measure else t
    x, y = xd.detach().requires_grad_(True), yd.detach().requires_grad_(True); t = tick("detach/requires_grad", t) if measure else t
    loss = sharded_chamfer_loss(x, y, total_batch=B);               t = tick("loss (autograd forward)", t) if measure else t
    rd.push(loss);                                                  t = tick("reader.push", t) if measure else t
    loss.backward();                                                t = tick("backward", t) if measure else t
    pf.release(); pf.prefetch((ah, bh));                            t = tick("release + prefetch", t) if measure else t
    if len(rd) > 1: rd.pop()
    t = tick("reader.pop", t) if measure else t
for _ in range(50): step(False)
torch.cuda.synchronize()
n = 400
t0 = time.perf_counter()
for _ in range(n): step(True)
torch.cuda.synchronize()
tot = (time.perf_counter() - t0) / n * 1e6
for k, v in acc.items(): print("%-28s %7.1f us" % (k, v / n * 1e6))
print("%-28s %7.1f us (wall per step)" % ("total", tot))
