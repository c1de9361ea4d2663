#!/usr/bin/env python
"""How fast is the library TF32 GEMM at the K/V-projection shape?  (context for the tcgen05 kernel's ceiling)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infinite_video_b200 import ops
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = True
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (M, N, K) in ((32768, 1536, 768), (8192, 8192, 8192), (32768, 1536, 4096)):
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); bias = torch.zeros(N, device=dev)
    out = torch.empty(M, N, device=dev)
    t_lib = timeit(lambda: torch.matmul(A, W.t(), out=out))
    t_own = timeit(lambda: ops.project_kv(A, W, bias, "tf32", out=out)) if N % 256 == 0 else float("nan")
    Ab, Wb = A.bfloat16(), W.bfloat16(); ob = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    t_bf = timeit(lambda: torch.matmul(Ab, Wb.t(), out=ob))
    fl = 2.0 * M * N * K
    print(f"M={M} N={N} K={K}: cublas tf32 {t_lib*1e3:.1f} us {fl/t_lib/1e9:.0f} TF/s | own tf32 {t_own*1e3:.1f} us "
          f"{fl/t_own/1e9:.0f} TF/s | cublas bf16 {t_bf*1e3:.1f} us {fl/t_bf/1e9:.0f} TF/s", flush=True)
# ---- fp16-operand path of the own kernel at the K/V projection shape
M, N, K = 32768, 1536, 768
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.03; bias = torch.zeros(N, device=dev)
Ah, Wh = A.half(), W.half(); out = torch.empty(M, N, device=dev)
t = timeit(lambda: ops.gemm_fp16(Ah, Wh, bias, out=out, round_tf32=True))
t1 = timeit(lambda: ops.project_kv_r(A, W, bias, out=out))
print(f"own fp16 operands {t*1e3:.1f} us ({2.0*M*N*K/t/1e9:.0f} TF/s) | own tf32 {t1*1e3:.1f} us")
