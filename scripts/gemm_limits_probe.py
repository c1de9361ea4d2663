#!/usr/bin/env python
"""Where does the tcgen05 GEMM lose time?  Times the kernel with the epilogue stores and / or the TMA loads removed
(bring-up flags, results are garbage) for the single-CTA and the CTA-pair variants."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infinite_video_b200 import _capi, ops
lib = _capi.lib()
for f in ("ltm_debug_set_pair", "ltm_debug_set_gemm_flags", "ltm_debug_set_cluster"):
    getattr(lib, f).argtypes = [C.c_int]; getattr(lib, f).restype = None
dev = torch.device("cuda:0")
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
names = {0: "full", 1: "no-store", 2: "no-tma", 3: "mma-only", 4: "no-STG"}
for (M, N, K) in ((32768, 1536, 768), (32768, 1536, 4096)):
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); bias = torch.zeros(N, device=dev)
    out = torch.empty(M, N, device=dev)
    fl = 2.0 * M * N * K
    for prec in ("tf32", "tf32x3"):
        for pair in (0, 1, 2):
            lib.ltm_debug_set_pair(1 if pair == 1 else 0)
            lib.ltm_debug_set_cluster(2 if pair == 2 else 1)
            line = f"M={M} N={N} K={K} {prec:7s} {('single', 'pair  ', 'mcast2')[pair]}:"
            for flags in (0, 1, 2, 3, 4):
                lib.ltm_debug_set_gemm_flags(flags)
                t = timeit(lambda: ops.project_kv(A, W, bias, prec, out=out))
                line += f"  {names[flags]} {t*1e3:7.1f} us ({fl/t/1e9:4.0f} TF/s)"
            lib.ltm_debug_set_gemm_flags(0)
            print(line, flush=True)
lib.ltm_debug_set_pair(0); lib.ltm_debug_set_cluster(1)
