#!/usr/bin/env python
"""Kernel durations inside the single-video CUDA-graph replay (CUPTI): how much of the ~35 us per call is kernels and
how much the gaps between five dependent launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from infinite_video_b200.batched import BatchedRectLTM
dev = torch.device("cuda:0")
E = D = 768
torch.manual_seed(0)
key, val = torch.nn.Linear(E, D), torch.nn.Linear(E, D)
eng = BatchedRectLTM(256, .75, key.weight.detach(), key.bias.detach(), val.weight.detach(), val.bias.detach(), device=dev)
k = torch.randn(1, 8192, E, device=dev); q = torch.randn(1, 32, D, device=dev)
u = torch.rand(1, 512, device=dev, dtype=torch.float64)
eng.step(k, q, None, new_doc=True)
side = torch.cuda.Stream(device=dev)
side.wait_stream(torch.cuda.current_stream(dev))
with torch.cuda.stream(side):
    eng.step(k, q, u); eng.step(k, q, u)
torch.cuda.current_stream(dev).wait_stream(side)
graphs = []
for _ in range(2):
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        eng.step(k, q, u)
    graphs.append(gr)
for i in range(10):
    graphs[i & 1].replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(20):
        graphs[i & 1].replay()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
for e in evs[50:60]:
    print(f"{e.time_range.start - t0:9.1f} us  +{e.time_range.end - e.time_range.start:6.1f} us  {e.name[:70]}")
tot = sum(e.time_range.end - e.time_range.start for e in evs)
print("kernel time per call", tot / 20, "us; span per call", (evs[-1].time_range.end - t0) / 20, "us")
