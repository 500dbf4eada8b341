"""Generates the golden vectors under tests/golden/ by running the REFERENCE's own CUDA
kernels (oracle/_ref/ref_losses.so, ref_sampling.so -- the unmodified reference sources
compiled for sm_100a by oracle/build_ref.sh) on a B200.

Run on the GPU box:   python tests/golden/make_golden.py gpurun_out/golden
then copy gpurun_out/golden/*.npz into tests/golden/ and commit them.  The vectors pin the CPU
oracle (tests/test_golden.py, CPU-only) and are re-checked against our kernels (-m gpu).
Inputs are stored next to the outputs, so the fixtures are self-contained.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

from helpers import lattice_cloud, sphere_cloud, uniform_cloud, with_duplicates  # noqa: E402


def ref_chamfer(rl, a, b):
    a, b = a.cuda().contiguous(), b.cuda().contiguous()
    B, N, _ = a.shape
    M = b.shape[1]
    d1 = torch.zeros(B, N, device="cuda"); d2 = torch.zeros(B, M, device="cuda")
    i1 = torch.zeros(B, N, dtype=torch.int32, device="cuda"); i2 = torch.zeros(B, M, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    rl.nmdistance_forward(a, b, d1, d2, i1, i2)   # launches on the legacy default stream
    torch.cuda.synchronize()
    return d1, d2, i1, i2


def ref_chamfer_bwd(rl, a, b, gd1, gd2, i1, i2):
    a, b = a.cuda().contiguous(), b.cuda().contiguous()
    g1, g2 = torch.zeros_like(a), torch.zeros_like(b)
    torch.cuda.synchronize()
    rl.nmdistance_backward(a, b, g1, g2, gd1.cuda().contiguous(), gd2.cuda().contiguous(), i1, i2)
    torch.cuda.synchronize()
    return g1, g2


def ref_labeled(rl, a, b, la, lb):
    a, b = a.cuda().contiguous(), b.cuda().contiguous()
    B, N, _ = a.shape
    M = b.shape[1]
    d1 = torch.zeros(B, N, device="cuda"); d2 = torch.zeros(B, M, device="cuda")
    i1 = torch.zeros(B, N, dtype=torch.int32, device="cuda"); i2 = torch.zeros(B, M, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    rl.labeled_nmdistance_forward(a, b, la.cuda().float(), lb.cuda().float(), d1, d2, i1, i2)
    torch.cuda.synchronize()
    return d1, d2, i1, i2


def ref_fps(rs, x, m, seed):
    x = x.cuda().contiguous()
    B, N, _ = x.shape
    idx = torch.empty(B, m, dtype=torch.int32, device="cuda")
    temp = torch.full((B, N), 1e10, device="cuda")
    torch.cuda.synchronize()
    rs.furthest_sampling(m, seed, x, temp, idx)
    torch.cuda.synchronize()
    return idx, temp


def n(t):
    return t.detach().cpu().numpy()


def main(out_dir):
    import ref_losses as rl
    import ref_sampling as rs
    os.makedirs(out_dir, exist_ok=True)
    # ---- chamfer forward/backward
    cases = {
        "chamfer_uniform": (uniform_cloud(2, 300, 101), uniform_cloud(2, 257, 102)),
        "chamfer_chunks": (uniform_cloud(2, 1100, 103), uniform_cloud(2, 1500, 104)),   # several 512-chunks
        "chamfer_sphere": (sphere_cloud(2, 700, 105), sphere_cloud(2, 600, 106)),
        "chamfer_ties": (with_duplicates(lattice_cloud(2, 900, 107, 6)), lattice_cloud(2, 1300, 108, 6)),
        "chamfer_dim5": (uniform_cloud(2, 200, 109, c=5), uniform_cloud(2, 333, 110, c=5)),
    }
    for name, (a, b) in cases.items():
        d1, d2, i1, i2 = ref_chamfer(rl, a, b)
        gd1 = uniform_cloud(a.shape[0], a.shape[1], 111, c=1)[..., 0]
        gd2 = uniform_cloud(b.shape[0], b.shape[1], 112, c=1)[..., 0]
        g1, g2 = ref_chamfer_bwd(rl, a, b, gd1, gd2, i1, i2)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), xyz1=n(a), xyz2=n(b), dist1=n(d1), dist2=n(d2),
                            idx1=n(i1), idx2=n(i2), gd1=n(gd1), gd2=n(gd2), g1=n(g1), g2=n(g2))
    # ---- labeled
    a, b = uniform_cloud(2, 700, 113), uniform_cloud(2, 900, 114)
    g = torch.Generator().manual_seed(115)
    la = torch.randint(0, 4, (2, 700, 1), generator=g); lb = torch.randint(0, 3, (2, 900, 1), generator=g)
    d1, d2, i1, i2 = ref_labeled(rl, a, b, la, lb)
    np.savez_compressed(os.path.join(out_dir, "chamfer_labeled.npz"), xyz1=n(a), xyz2=n(b), label1=n(la), label2=n(lb),
                        dist1=n(d1), dist2=n(d2), idx1=n(i1), idx2=n(i2))
    # ---- FPS (ties included), gather
    fps_cases = {
        "fps_small": (uniform_cloud(2, 100, 120), 30, 0),
        "fps_511": (uniform_cloud(2, 511, 121), 64, 7),
        "fps_2k": (uniform_cloud(2, 2000, 122), 200, 3),
        "fps_sphere": (sphere_cloud(2, 3000, 123), 128, 0),
        "fps_dups": (with_duplicates(uniform_cloud(2, 4096, 124), 0.3), 300, 1),
        "fps_lattice": (lattice_cloud(2, 4096, 125, 6), 300, 1),
        "fps_16k": (uniform_cloud(1, 16384, 126), 256, 0),
    }
    for name, (x, m, seed) in fps_cases.items():
        idx, temp = ref_fps(rs, x, m, seed)
        feats = x.transpose(1, 2).contiguous().cuda()
        out = torch.empty(x.shape[0], 3, m, device="cuda")
        rs.gather_forward(x.shape[0], 3, x.shape[1], m, feats, idx, out)
        gout = uniform_cloud(x.shape[0], m, 127, c=3).transpose(1, 2).contiguous().cuda()
        gin = torch.zeros(x.shape[0], 3, x.shape[1], device="cuda")
        rs.gather_backward(x.shape[0], 3, x.shape[1], m, gout, idx, gin)
        torch.cuda.synchronize()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), xyz=n(x), m=m, seed=seed, idx=n(idx), temp=n(temp),
                            gathered=n(out), grad_out=n(gout), grad_in=n(gin))
    # ---- ball_query + group
    bq_cases = {
        "bq_uniform": (uniform_cloud(2, 4096, 130), uniform_cloud(2, 256, 131), 0.2, 32),
        "bq_sparse": (uniform_cloud(2, 1000, 132), uniform_cloud(2, 200, 133), 0.05, 16),
        "bq_sphere": (sphere_cloud(2, 3000, 134), sphere_cloud(2, 100, 135), 0.2, 32),
        "bq_lattice": (lattice_cloud(2, 777, 136, 8), lattice_cloud(2, 33, 137, 8), 0.25, 8),
        "bq_big_nsample": (uniform_cloud(1, 10, 138), uniform_cloud(1, 4, 139), 10.0, 64),
    }
    for name, (xyz, ctr, r, ns) in bq_cases.items():
        idx = rs.ball_query(ctr.cuda().contiguous(), xyz.cuda().contiguous(), r, ns)
        feats = xyz.transpose(1, 2).contiguous().cuda()
        grouped = rs.group_points(feats, idx)
        ggrad = rs.group_points_grad(torch.ones_like(grouped), idx, xyz.shape[1])
        torch.cuda.synchronize()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), xyz=n(xyz), new_xyz=n(ctr), radius=r, nsample=ns,
                            idx=n(idx), grouped=n(grouped), group_grad=n(ggrad))
    # ---- three_nn
    u, k = uniform_cloud(2, 777, 140), uniform_cloud(2, 1300, 141)
    d = torch.empty(2, 777, 3, device="cuda"); i = torch.empty(2, 777, 3, dtype=torch.int32, device="cuda")
    rs.three_nn_wrapper(2, 777, 1300, u.cuda().contiguous(), k.cuda().contiguous(), d, i)
    torch.cuda.synchronize()
    np.savez_compressed(os.path.join(out_dir, "three_nn.npz"), unknown=n(u), known=n(k), dist2=n(d), idx=n(i))
    # ---- three_interpolate (weights as pointnet2_modules.py:131-135 builds them)
    dist = torch.sqrt(d)
    w = 1.0 / (dist + 1e-8)
    w = (w / w.sum(dim=2, keepdim=True)).contiguous()
    feats = uniform_cloud(2, 1300, 142, c=6).transpose(1, 2).contiguous().cuda()
    out = torch.empty(2, 6, 777, device="cuda")
    rs.three_interpolate_wrapper(2, 6, 1300, 777, feats, i, w, out)
    gout = uniform_cloud(2, 777, 143, c=6).transpose(1, 2).contiguous().cuda()
    gin = torch.zeros(2, 6, 1300, device="cuda")
    rs.three_interpolate_grad_wrapper(2, 6, 777, 1300, gout, i, w, gin)
    torch.cuda.synchronize()
    np.savez_compressed(os.path.join(out_dir, "three_interpolate.npz"), features=n(feats), idx=n(i), weight=n(w),
                        out=n(out), grad_out=n(gout), grad_in=n(gin))
    print("golden vectors written to", out_dir, sorted(os.listdir(out_dir)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
