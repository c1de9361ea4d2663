#!/usr/bin/env python
"""GPU probe: Gaussian-variant throughput vs batch, with per-op device times."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda:0")
for Bv, Lk in ((32, 256), (128, 256), (256, 256), (32, 8192)):
    r = bench.run_gauss_arm(dev, Bv, 3, Lk)
    print(Bv, Lk, round(r["value"]), "chunks/s", round(r["ms_per_step"], 3), "ms/step", round(r["roofline"]["frac"], 3))
# per-op timing at Bv=128
from infinite_video_b200 import ops
from infinite_video_b200.batched import BatchedGaussLTM
torch.manual_seed(0)
key, val = torch.nn.Linear(768, 768), torch.nn.Linear(768, 768)
eng = BatchedGaussLTM(256, .75, key.weight.detach(), key.bias.detach(), val.weight.detach(), val.bias.detach(), device=dev)
Bv, L = 128, 256
k = torch.randn(Bv, L, 768, device=dev); q = torch.randn(Bv, 32, 768, device=dev)
u = torch.rand(Bv, 512, device=dev, dtype=torch.float64)
eng.step(k, q, None, True); eng.step(k, q, u, False)
op = eng.operators(L)
def timeit(name, fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = fn()
    e1.record(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): out = fn()
    cpu = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    print(f"  {name:28s} gpu {e0.elapsed_time(e1)/n:7.3f} ms   host-enqueue {cpu:6.3f} ms")
    return out
hist = timeit("sticky_hist_gauss", lambda: ops.sticky_hist_gauss(eng._mu, eng._sd, op["tb"]))
rs = timeit("resample+sort", lambda: ops.resample(hist, u, op["bins"], None, normalize=True, sort=True))
R = timeit("gemm Psi_tab @ B_past (x3)", lambda: ops.gemm(op["Psi_tab"], eng._B, a_kmajor=True, b_kmajor=False, precision="tf32x3"))
xm = timeit("gather_rows", lambda: ops.gather_rows(R, rs["b_used"]))
B = timeit("gemm G_inf^T [xm;k] (x3)", lambda: ops.gemm(op["GinfT"], xm, B2=k, a_kmajor=True, b_kmajor=False, precision="tf32x3"))
timeit("gemm G_inf^T [xm;k] (x1)", lambda: ops.gemm(op["GinfT"], xm, B2=k, a_kmajor=True, b_kmajor=False, precision="tf32"))
Kt, V = timeit("project_kv_t (x3)", lambda: ops.project_kv_t(B, eng.Wkv, eng.bkv, 256, precision="tf32x3"))
timeit("project_kv_t (x1)", lambda: ops.project_kv_t(B, eng.Wkv, eng.bkv, 256, precision="tf32"))
timeit("cont_attn_gauss_t", lambda: ops.cont_attn_gauss_t(q, Kt, V, op["mu"], op["sigma"]))
timeit("whole step", lambda: eng.step(k, q, u, False))
