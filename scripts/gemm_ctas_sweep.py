#!/usr/bin/env python
"""Overlapped chunk loop at the BASELINE shapes with the K/V projection confined to n CTAs while the next chunk is pooled
beside the step (0 = one CTA per SM; default = the engine's choice), same box, same process."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from infinite_video_b200.batched import BatchedRectLTM
dev = torch.device("cuda:0")
D, S, TAU = 768, 512, 0.75
shapes = dict(bench.OTHER_CONFIGS)
shapes["cfg2"] = dict(videos=128, N=256, L=256, T=32, e=768, Q=32)
for name in ("cfg1", "cfg3", "cfg4", "cfg2"):
    cfg = shapes[name]
    Bv, N, Lc, Tc, e, Qc = cfg["videos"], cfg["N"], cfg["L"], cfg["T"], cfg["e"], cfg["Q"]
    torch.manual_seed(0)
    key, val = torch.nn.Linear(e, D), torch.nn.Linear(e, D)
    g = torch.Generator(device=dev).manual_seed(77)
    C = 8
    ks = [torch.randn(Bv, Lc * Tc, e, device=dev, generator=g) for _ in range(C)]
    qs = [torch.randn(Bv, Qc, D, device=dev, generator=g) for _ in range(C)]
    us = [torch.rand(Bv, S, device=dev, dtype=torch.float64, generator=g) for _ in range(C)]
    for ctas in (None, 0, 32, 64, 96, 0, None):
        eng = BatchedRectLTM(N, TAU, key.weight.detach(), key.bias.detach(), val.weight.detach(), val.bias.detach(),
                             tokens_per_frame=Tc, sticky=True, device=dev)
        if ctas is not None:
            eng.gemm_ctas_overlap = ctas
            eng.gemm_ctas_min_ratio = 0
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
        print(f"{name} projection CTAs beside pooling = {ctas}: {Bv * C / (ms * 1e-3):.0f} chunks/s", flush=True)
        del eng
    del ks, qs, us
    torch.cuda.empty_cache()
