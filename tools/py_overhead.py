"""Host-side cost per call of the layers of the reference-signature Chamfer path (B=32, N=M=2500: the GPU step
is ~0.09 ms, so the Python above it decides the end-to-end rate).  Prints microseconds of HOST time per call
(the GPU runs behind; a synchronize every 50 calls keeps the queue bounded)."""
import sys, os, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import uniform_cloud
from pytorch_points_b200 import _C
from pytorch_points_b200._ext import losses
from pytorch_points_b200.network import model_loss
from pytorch_points_b200.dist import sharded_chamfer_loss
from pytorch_points_b200.pipeline import HostPrefetcher
B, N = 32, 2500
dev = torch.device("cuda")
ah, bh = uniform_cloud(B, N, 1).pin_memory(), uniform_cloud(B, N, 2).pin_memory()
a, b = ah.to(dev), bh.to(dev)
d1 = torch.empty(B, N, device=dev); d2 = torch.empty(B, N, device=dev)
i1 = torch.empty(B, N, dtype=torch.int32, device=dev); i2 = torch.empty(B, N, dtype=torch.int32, device=dev)
sums = torch.zeros(2, device=dev); gw = torch.full((2,), 1.0 / (B * N), device=dev)
g1, g2 = torch.empty_like(a), torch.empty_like(b)


def host_us(fn, n=400):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(n):
        fn()
        if k % 50 == 49: torch.cuda.synchronize()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def ag_nndistance():
    x, y = a.detach().requires_grad_(True), b.detach().requires_grad_(True)
    o1, o2, _, _ = model_loss.nndistance(x, y)
    (o1.mean() + o2.mean()).backward()


def ag_sums():
    x, y = a.detach().requires_grad_(True), b.detach().requires_grad_(True)
    sharded_chamfer_loss(x, y, total_batch=B).backward()


def ag_mean():
    x, y = a.detach().requires_grad_(True), b.detach().requires_grad_(True)
    model_loss.chamfer_mean_loss(x, y).backward()

pf = HostPrefetcher(dev, depth=2); pf.prefetch((ah, bh))
def prefetch_cycle():
    pf.get(); pf.release(); pf.prefetch((ah, bh))

print("5 x torch.empty                         %7.1f us" % host_us(lambda: [torch.empty(B, N, device=dev) for _ in range(5)]))
print("_ext fused fwd+bwd call                 %7.1f us" % host_us(lambda: losses.nmdistance_forward_backward_uniform(a, b, d1, d2, i1, i2, sums, gw, g1, g2)))
print("_ext forward call                       %7.1f us" % host_us(lambda: losses.nmdistance_forward(a, b, d1, d2, i1, i2, sums=sums)))
print("prefetcher get/release/prefetch         %7.1f us" % host_us(prefetch_cycle))
print("autograd nndistance + means + backward  %7.1f us" % host_us(ag_nndistance))
print("autograd sharded_chamfer_loss + backward%7.1f us" % host_us(ag_sums))
print("autograd chamfer_mean_loss + backward   %7.1f us" % host_us(ag_mean))
if "profile" in sys.argv:
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for _ in range(300): ag_mean()
    torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
