#!/usr/bin/env python
"""Stage times (serial pass, CUDA events around each kernel) of the rect path at the BASELINE shapes."""
import ctypes as Ct, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infinite_video_b200 import _capi
from infinite_video_b200.batched import BatchedRectLTM
dev = torch.device("cuda:0")
lib = _capi.lib()
names = ["pool", "resample", "consolidate", "project_kv", "attention"]
torch.manual_seed(0)
for name, (Bv, N, L, T, e, Q) in {"cfg1": (1024, 64, 8, 32, 768, 32), "cfg3": (128, 64, 16, 196, 1024, 96),
                                  "cfg4": (64, 512, 256, 32, 768, 32), "cfg2": (128, 256, 256, 32, 768, 32)}.items():
    key, val = torch.nn.Linear(e, 768), torch.nn.Linear(e, 768)
    eng = BatchedRectLTM(N, .75, key.weight.detach(), key.bias.detach(), val.weight.detach(), val.bias.detach(),
                         tokens_per_frame=T, device=dev)
    ks = [torch.randn(Bv, L * T, e, device=dev) for _ in range(4)]
    q = torch.randn(Bv, Q, 768, device=dev)
    u = torch.rand(Bv, 512, dtype=torch.float64, device=dev)
    eng.step(ks[0], q, None, new_doc=True)
    for i in range(3):
        eng.step(ks[i], q, u)
    evs = []
    for _ in range(10):
        h = Ct.c_void_p()
        _capi.check(lib.ltm_event_create(Ct.byref(h)), "event_create")
        evs.append(h)
    acc = {n: 0.0 for n in names}
    reps = 8
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(reps):
        eng.prof_events = evs
        eng.step(ks[i % 4], q, u)
        torch.cuda.synchronize()
        f = Ct.c_float()
        for j, n in enumerate(names):
            if lib.ltm_event_elapsed_ms(evs[2 * j], evs[2 * j + 1], Ct.byref(f)) == 0:
                acc[n] += f.value / reps
    eng.prof_events = None
    gb = Bv * (L * T * e + 2 * Q * 768 + 2 * N * e) * 4 / 1e9
    tot = sum(acc.values())
    print(f"{name}: Bv={Bv} N={N} L={L} T={T} e={e} Q={Q} kv_state={eng.kv_state} tc={eng.tc_attn}: "
          + " ".join(f"{n}={v*1e3:.0f}us" for n, v in acc.items())
          + f" | sum {tot*1e3:.0f}us, roofline {gb/6545*1e3*1e3:.0f}us -> {gb/6545*1e3/tot:.2f}", flush=True)
    del eng, ks
    torch.cuda.empty_cache()
