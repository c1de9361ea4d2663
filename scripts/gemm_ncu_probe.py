#!/usr/bin/env python
"""Launches the K/V projection GEMM in each variant once (for `ncu -k regex:gemm_tf32`)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infinite_video_b200 import _capi, ops
lib = _capi.lib()
lib.ltm_debug_set_pair.argtypes = [C.c_int]; lib.ltm_debug_set_pair.restype = None
dev = torch.device("cuda:0")
A = torch.randn(32768, 768, device=dev); B = torch.randn(1536, 768, device=dev); bias = torch.zeros(1536, device=dev)
out = torch.empty(32768, 1536, device=dev)
for pair in (0, 1):
    lib.ltm_debug_set_pair(pair)
    for prec in ("tf32", "tf32x3"):
        for _ in range(2):
            ops.project_kv(A, B, bias, prec, out=out)
torch.cuda.synchronize()
