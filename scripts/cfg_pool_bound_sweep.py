#!/usr/bin/env python
"""Overlapped chunk loop at the smaller BASELINE shapes with the side-stream pooling grid bounded to n CTAs per SM
(0 = unbounded): when the chain of the other stream is as long as the pooling, leaving it room may pay."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from infinite_video_b200.batched import BatchedRectLTM
dev = torch.device("cuda:0")
D, S, TAU = 768, 512, 0.75
for name in ("cfg3", "cfg4", "cfg1"):
    cfg = bench.OTHER_CONFIGS[name]
    Bv, N, Lc, Tc, e, Qc = cfg["videos"], cfg["N"], cfg["L"], cfg["T"], cfg["e"], cfg["Q"]
    torch.manual_seed(0)
    key, val = torch.nn.Linear(e, D), torch.nn.Linear(e, D)
    g = torch.Generator(device=dev).manual_seed(77)
    C = 8
    ks = [torch.randn(Bv, Lc * Tc, e, device=dev, generator=g) for _ in range(C)]
    qs = [torch.randn(Bv, Qc, D, device=dev, generator=g) for _ in range(C)]
    us = [torch.rand(Bv, S, device=dev, dtype=torch.float64, generator=g) for _ in range(C)]
    for bound, binp in ((0, None), (0, False), (3, False), (4, False), (5, False), (6, False)):
        eng = BatchedRectLTM(N, TAU, key.weight.detach(), key.bias.detach(), val.weight.detach(), val.bias.detach(),
                             tokens_per_frame=Tc, sticky=True, device=dev, bin_pool=binp)
        eng.pool_ctas = bound * 148
        def one_step():
            for c in range(C):
                eng.step_overlapped(ks[c], qs[c], us[c] if c else None, new_doc=(c == 0), k_next=ks[(c + 1) % C],
                                    next_new_doc=((c + 1) % C == 0))
        for _ in range(3):
            one_step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(5):
            one_step()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / 5
        print(f"{name} pool_ctas/SM={bound} bin_pool={binp}: {ms / C * 1e3:.0f} us per chunk-step, {Bv * C / (ms * 1e-3):.0f} chunks/s",
              flush=True)
        del eng
    del ks, qs, us
    torch.cuda.empty_cache()
