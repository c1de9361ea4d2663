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


def compare_draws(b_got, b_want, u, p, delta):
    """Sampled bins of the implementation under test vs the oracle's own draws from the same uniforms, NO guard band.
    Returns (flips, draws).  A draw may only differ where it is a numerical tie: every CDF edge of the oracle's `p`
    (fp32 [B,C]) that separates the two bins must lie within `delta` of the uniform -- i.e. the two histograms
    differ by less than `delta` there.  Anything else raises."""
    b_got, b_want = torch.as_tensor(b_got).long().cpu(), torch.as_tensor(b_want).long().cpu()
    diff = b_got != b_want
    flips = int(diff.sum())
    if flips:
        p64 = torch.as_tensor(p).double().cpu()
        cdf = torch.cumsum(p64, -1) / p64.sum(-1, keepdim=True)
        for v, s in diff.nonzero().tolist():
            lo, hi = sorted((int(b_got[v, s]), int(b_want[v, s])))
            uu = float(u[v, s])
            assert cdf[v, lo] > uu - delta and cdf[v, hi - 1] < uu + delta, \
                f"draw {s} of video {v}: bins {int(b_got[v, s])} vs {int(b_want[v, s])} is not a tie within {delta}"
    return flips, b_got.numel()


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
