#!/usr/bin/env python
"""Launch list for ncu: each bring-up variant of the K/V-projection GEMM and the library TF32 GEMM, 3x each.
ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infinite_video_b200 import _capi, ops
lib = _capi.lib()
for f in ("ltm_debug_set_pair", "ltm_debug_set_gemm_flags"):
    getattr(lib, f).argtypes = [C.c_int]; getattr(lib, f).restype = None
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = True
M, N, K = 32768, 1536, 768
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); bias = torch.zeros(N, device=dev)
out = torch.empty(M, N, device=dev)
for pair in (0, 1):
    lib.ltm_debug_set_pair(pair)
    for flags in (0, 1, 2, 3, 4):
        lib.ltm_debug_set_gemm_flags(flags)
        for _ in range(3):
            ops.project_kv(A, W, bias, "tf32", out=out)
lib.ltm_debug_set_gemm_flags(0); lib.ltm_debug_set_pair(0)
for _ in range(3):
    torch.matmul(A, W.t(), out=out)
torch.cuda.synchronize()
