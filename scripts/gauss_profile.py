#!/usr/bin/env python
"""Kernel-level timeline of variant G (CUPTI kernel times of update calls at 128 videos, 256 un-pooled frames)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from infinite_video_b200.batched import BatchedGaussLTM
dev = torch.device("cuda:0")
Bv, L, E, Q, N = int(os.environ.get("BV", 128)), 256, 768, 32, 256
torch.manual_seed(0)
key, val = torch.nn.Linear(E, 768), torch.nn.Linear(E, 768)
eng = BatchedGaussLTM(N, .75, key.weight.detach(), key.bias.detach(), val.weight.detach(), val.bias.detach(), device=dev,
                      proj_precision=os.environ.get("PROJ") or None)
ks = [torch.randn(Bv, L, E, device=dev) for _ in range(3)]
qs = [torch.randn(Bv, Q, 768, device=dev) for _ in range(3)]
us = [torch.rand(Bv, 512, dtype=torch.float64, device=dev) for _ in range(3)]
def one():
    for c in range(3):
        eng.step(ks[c], qs[c], us[c] if c else None, new_doc=(c == 0))
one(); one()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    one()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=20, max_name_column_width=90))
