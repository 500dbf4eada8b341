"""Shared input generators for the parity tests (seeded, CPU-generated, identical bits on
every path -- SURVEY.md §8d)."""
import math

import numpy as np
import torch


def uniform_cloud(B, N, seed, c=3):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, N, c, generator=g, dtype=torch.float32)


def sphere_cloud(B, N, seed):
    """Same construction as the reference's utils/pc_utils.py:504-516 random_sphere
    (theta = 2*pi*u, phi = pi*v => pole-clustered), float64 -> float32."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(B, N, generator=g, dtype=torch.float64)
    v = torch.rand(B, N, generator=g, dtype=torch.float64)
    theta = 2 * math.pi * u
    phi = math.pi * v
    x = torch.cos(theta) * torch.sin(phi)
    y = torch.sin(theta) * torch.sin(phi)
    z = torch.cos(phi)
    return torch.stack([x, y, z], dim=-1).to(torch.float32)


def with_duplicates(x, frac=0.1):
    """Last `frac` of the points duplicated from the first `frac` (the padding the reference's
    own loaders produce, utils/pc_utils.py:222-227) -> exact distance ties."""
    x = x.clone()
    n = x.shape[1]
    k = max(1, int(n * frac))
    x[:, n - k:] = x[:, :k]
    return x


def lattice_cloud(B, N, seed, levels=8):
    """Points on a coarse lattice: many exactly equal distances -> stresses every tie-break."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, levels, (B, N, 3), generator=g).to(torch.float32) / levels


def np32(t):
    return np.ascontiguousarray(t.detach().cpu().numpy())
