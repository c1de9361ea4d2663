#!/usr/bin/env python
"""Serial step time of the rect path at the other BASELINE shapes (cfg1, cfg3, cfg4) -- parity-test configs, timed
here only to see which kernels they land on."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infinite_video_b200.batched import BatchedRectLTM
dev = torch.device("cuda:0")
torch.manual_seed(0)
for name, (Bv, N, L, T, e, Q) in {"cfg1": (128, 64, 8, 32, 768, 32), "cfg3": (64, 64, 16, 196, 1024, 96),
                                  "cfg4": (64, 512, 256, 32, 768, 32), "cfg2": (128, 256, 256, 32, 768, 32)}.items():
    key, val = torch.nn.Linear(e, 768), torch.nn.Linear(e, 768)
    eng = BatchedRectLTM(N, .75, key.weight.detach(), key.bias.detach(), val.weight.detach(), val.bias.detach(),
                         tokens_per_frame=T, device=dev)
    k = torch.randn(Bv, L * T, e, device=dev); q = torch.randn(Bv, Q, 768, device=dev)
    u = torch.rand(Bv, 512, dtype=torch.float64, device=dev)
    eng.step(k, q, None, new_doc=True)
    for _ in range(3): eng.step(k, q, u)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): eng.step(k, q, u)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = Bv * (L * T * e + 2 * Q * 768 + 2 * N * e) * 4 / 1e9
    print(f"{name}: Bv={Bv} N={N} L={L} T={T} e={e} Q={Q}: {ms:.3f} ms/step, {Bv/ms*1e3:.0f} chunks/s, "
          f"{gb/ms*1e3/6545:.2f} of the HBM roofline, tc_attn={eng.tc_attn}", flush=True)
    del eng, k, q
