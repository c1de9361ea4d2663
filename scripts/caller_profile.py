#!/usr/bin/env python
"""Kernel-level timeline of the caller branch (N1): CUPTI kernel times of one call at 32 videos."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from infinite_video_b200.cross_attention import CrossAttentionLTM
dev = torch.device("cuda:0")
Bv, L, T, E, Q, D, H, NB = int(os.environ.get("BV", 32)), 256, 32, 768, 32, 768, 12, 256
torch.manual_seed(0)
lin = lambda: torch.nn.Linear(E, D).to(dev)
mod = CrossAttentionLTM(lin(), lin(), lin(), alpha=0.9, num_basis=NB, tau=.75, sticky=True, n_heads=H)
ks = [torch.randn(Bv, L * T, E, device=dev) for _ in range(3)]
hs = [torch.randn(Bv, Q, D, device=dev) for _ in range(3)]
us = [torch.rand(Bv, 512, dtype=torch.float64, device=dev) for _ in range(3)]
def one():
    for c in range(3):
        out = mod(hs[c], ks[c], new_video=(c == 0), u=us[c] if c else None)
    return out
one(); one()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    one()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
