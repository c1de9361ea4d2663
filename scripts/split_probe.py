#!/usr/bin/env python
"""GPU probe: does tcgen05 kind::tf32 truncate fp32 operands (then the splitter need not rewrite `hi`)?"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infinite_video_b200 import _capi, ops
lib = _capi.lib()
class _Noop:
    def __call__(self, *a): pass
lib_set = _Noop()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
A = torch.randn(512, 768, generator=g); B = torch.randn(1536, 768, generator=g)
want = A.double() @ B.double().t()
for mode in (1, 0):
    lib_set(mode)
    got = ops.gemm(A.to(dev), B.to(dev), precision="tf32x3")[0].cpu().double()
    print("write_hi" if mode else "lo_only ", float((got - want).abs().max() / want.abs().max()))
lib_set(1)
lib.ltm_debug_set_pair.argtypes = [C.c_int]; lib.ltm_debug_set_pair.restype = None
for pair in (0, 1):
    lib.ltm_debug_set_pair(pair)
    for prec in ("tf32", "tf32x3"):
        got = ops.gemm(A.to(dev), B.to(dev), precision=prec)[0].cpu().double()
        print("pair" if pair else "single", prec, "relerr", float((got - want).abs().max() / want.abs().max()))
import time
for prec in ("tf32", "tf32x3"):
    for mode in (1, 0):
        lib.ltm_debug_set_pair(mode)
        Ab = torch.randn(32768, 768, device=dev); Bb = B.to(dev); bias = torch.zeros(1536, device=dev)
        out = torch.empty(32768, 1536, device=dev)
        ops.project_kv(Ab, Bb, bias, prec, out=out); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.project_kv(Ab, Bb, bias, prec, out=out)
        e1.record(); torch.cuda.synchronize()
        print(prec, "pair" if mode else "single", round(e0.elapsed_time(e1) / 10, 4), "ms")
lib_set(1)
