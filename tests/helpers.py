"""Shared test helpers: seeded inputs (identical to tests/golden/make_golden.py), tolerances, guard bands."""
import os

import numpy as np
import torch
import torch.nn as nn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# parity tolerances (BASELINE.json north_star): coefficients and context vectors within 1e-3 relative
# (max-abs error / max-abs reference) for the fp32-accumulate path; sampled bins bit-exact.
TOL_B = 1e-3
TOL_CTX = 1e-3


def make_inputs(seed, C, Bv, rows, e, Q, q_scale=1.0):
    g = torch.Generator().manual_seed(seed)
    ks = [torch.randn(Bv, rows, e, generator=g) for _ in range(C)]
    qs = [torch.randn(Bv, Q, 768, generator=g) * q_scale for _ in range(C)]
    us = [torch.rand(Bv, 512, dtype=torch.float64, generator=g) for _ in range(C)]
    return ks, qs, us


def make_proj(seed, e):
    torch.manual_seed(seed)
    return nn.Linear(e, 768), nn.Linear(e, 768)


def proj_tensors(key, val):
    return key.weight.detach(), key.bias.detach(), val.weight.detach(), val.bias.detach()


def relerr(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: z[k] for k in z.files}


def guard_band(u, p, eps=2e-6):
    """Move uniforms that sit within `eps` of a CDF edge of p (fp32 [B,C]) to the middle of their bin, so
    ulp-level differences between two correct implementations of p cannot flip a draw.  Returns a copy."""
    u = u.clone()
    p64 = p.double()
    cdf = torch.cumsum(p64, -1) / p64.sum(-1, keepdim=True)
    for i in range(u.shape[0]):
        d = (u[i].unsqueeze(1) - cdf[i].unsqueeze(0)).abs()
        close = d.min(1).values < eps
        if close.any():
            edges = torch.cat([torch.zeros(1, dtype=torch.float64), cdf[i]])
            mids = (edges[:-1] + edges[1:]) / 2
            big = (edges[1:] - edges[:-1]).argmax()
            u[i][close] = mids[big]
    return u
