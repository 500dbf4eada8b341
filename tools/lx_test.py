"""torchrun --nproc-per-node W tools/lx_test.py: LossExchange vs NCCL all-reduce (values + latency)."""
import os, sys, time, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_points_b200.dist import LossExchange
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
lx = LossExchange(dev)
sums = torch.zeros(2, device=dev); tot = torch.zeros(2, device=dev); ref = torch.zeros(2, device=dev)
ok = True
for step in range(50):
    sums.copy_(torch.tensor([rank + 1.0 + step, 0.25 * (rank + 1) * (step + 1)], device=dev))
    lx.send(sums)
    lx.wait(tot)
    ref.copy_(sums); dist.all_reduce(ref)
    torch.cuda.synchronize()
    if not torch.allclose(tot, ref, rtol=1e-6):
        ok = False; print("rank", rank, "step", step, tot.tolist(), ref.tolist(), flush=True)
# graph capture
g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
with torch.cuda.stream(s):
    lx.send(sums); lx.wait(tot); torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        lx.send(sums); lx.wait(tot)
dist.barrier(); torch.cuda.synchronize()
for step in range(20):
    sums.copy_(torch.tensor([rank + 3.0 + step, 7.0], device=dev)); torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    want0 = sum(r + 3.0 + step for r in range(world))
    if abs(tot[0].item() - want0) > 1e-3 or abs(tot[1].item() - 7.0 * world) > 1e-3:
        ok = False; print("graph rank", rank, step, tot.tolist(), want0, flush=True)
def bench(fn, n=200):
    for _ in range(20): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
t_lx = bench(lambda: (lx.send(sums), lx.wait(tot)))
t_nccl = bench(lambda: dist.all_reduce(ref))
t_g = bench(lambda: g.replay())
print("rank %d ok=%s timed_out=%s  peer exchange %.1f us/step (graph %.1f)  nccl all_reduce %.1f us/step" % (rank, ok, lx.timed_out(), t_lx, t_g, t_nccl), flush=True)
dist.destroy_process_group()
