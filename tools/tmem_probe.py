"""Tensor-memory read rate on this GPU (pp_microbench 7 / 8): the denominator of the sweep kernel's roofline."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from pytorch_points_b200 import _C
clk = torch.cuda.clock_rate() if hasattr(torch.cuda, "clock_rate") else None
for which, name in ((7, "tcgen05.ld only"), (8, "tcgen05.ld + granule minimum tree")):
    for iters in (2000, 20000):
        ms, b = _C.microbench(which, iters, 0)
        print("%-36s iters %6d: %.3f ms, %.2f TB/s chip, %.1f B/clk/SMSP at 1.965 GHz" % (
            name, iters, ms, b / ms / 1e9, b / (ms * 1e-3) / (148 * 4) / 1.965e9), flush=True)
for which, name in ((9, "FMNMX3 (in-place chains)"), (10, "FMNMX (two inputs)"), (11, "FMNMX3 (rotating sources)")):
    ms, n = _C.microbench(which, 20000, 0)
    print("%-36s %.3f ms: %.2f cycles per warp instruction per SMSP at 1.965 GHz" % (name, ms, ms * 1e-3 * 1.965e9 / (n / (148 * 4))), flush=True)
