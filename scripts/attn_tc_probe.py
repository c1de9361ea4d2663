#!/usr/bin/env python
"""Tensor-core attention vs the FMA fast path on the bench shape: error and time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infinite_video_b200 import ops, tables
dev = torch.device("cuda:0")
torch.manual_seed(0)
Bv, N, Q, D, H, L = int(sys.argv[1]) if len(sys.argv) > 1 else 128, 256, 32, 768, 12, 256
tab = tables.rect_tables(L, N, .75); td = tab.to(dev)
Bc = torch.randn(Bv * N, 768, device=dev)
Wkv = torch.randn(2 * D, 768, device=dev) * 0.03; bkv = torch.randn(2 * D, device=dev) * 0.1
for qs in (1.0, 4.0):
    q = torch.randn(Bv, Q, D, device=dev) * qs
    KVr = ops.project_kv_r(Bc, Wkv, bkv).view(Bv, N, 2 * D)
    KV = ops.project_kv(Bc, Wkv, bkv, precision="tf32x3").view(Bv, N, 2 * D)
    Kt = KV[:, :, :D].reshape(Bv, N, H, 64).permute(0, 2, 3, 1).contiguous(); V = KV[:, :, D:].contiguous()
    c0, s0, h0 = ops.cont_attn_rect_t(q, Kt, V, td["W"], tab.W_out, td["jb"], td["tb"], want_scores=True)
    c1, s1, h1 = ops.cont_attn_rect_tc(q, KVr, td["X"], td["W"], tab.W_out, tab.c_none, td["jb"], td["tb"], want_scores=True)
    rel = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()
    # the same rounded K|V through the FMA kernel: isolates the attention kernel's own error
    Ktr = KVr[:, :, :D].reshape(Bv, N, H, 64).permute(0, 2, 3, 1).contiguous(); Vr = KVr[:, :, D:].contiguous()
    c2, s2, h2 = ops.cont_attn_rect_t(q, Ktr, Vr, td["W"], tab.W_out, td["jb"], td["tb"], want_scores=True)
    print(f"   attention only (same K|V): scores {rel(s1, s2):.2e} ctx {rel(c1, c2):.2e} hist {rel(h1, h2):.2e}"
          f" | single-pass projection + FMA attention vs x3: scores {rel(s2, s0):.2e} ctx {rel(c2, c0):.2e} hist {rel(h2, h0):.2e}")
    print(f"q_scale {qs}: |S|max {s0.abs().max().item():.1f} scores {rel(s1, s0):.2e} ctx {rel(c1, c0):.2e} hist {rel(h1, h0):.2e}")
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("fma  ms", timeit(lambda: ops.cont_attn_rect_t(q, Kt, V, td["W"], tab.W_out, td["jb"], td["tb"])))
print("tc   ms", timeit(lambda: ops.cont_attn_rect_tc(q, KVr, td["X"], td["W"], tab.W_out, tab.c_none, td["jb"], td["tb"])))
# ---- per-item timeline of CTA 0 (bring-up trace)
import ctypes as C
from infinite_video_b200 import _capi
lib = _capi.lib()
lib.ltm_debug_set_attn_trace.argtypes = [C.c_void_p]; lib.ltm_debug_set_attn_trace.restype = None
tr = torch.zeros(16 * 16, dtype=torch.int64, device=dev)
lib.ltm_debug_set_attn_trace(C.c_void_p(tr.data_ptr()))
ops.cont_attn_rect_tc(q, KVr, td["X"], td["W"], tab.W_out, tab.c_none, td["jb"], td["tb"])
torch.cuda.synchronize()
lib.ltm_debug_set_attn_trace(None)
t = tr.cpu().view(16, 16).numpy()
names = {5: "S issue (K,q ready)", 8: "CW sees S", 0: "IS sees S done", 9: "maxima exchanged", 10: "e^T written", 1: "IS sees e^T", 2: "V ready", 3: "PV issued", 11: "CW sees PV", 4: "IS sees PV", 12: "normalisers", 13: "item end"}
t0 = t[0][5]
for it in range(min(8, Bv * 12 // 148 + 1)):
    ev = sorted((int(t[it][k] - t0), names[k]) for k in names if t[it][k] > 0)
    print(f"item {it}: " + " | ".join(f"{n} {dt/1000:.2f}" for dt, n in ev))
