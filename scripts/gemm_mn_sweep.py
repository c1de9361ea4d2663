#!/usr/bin/env python
"""Bring-up aid (GPU): sweep MN-major UMMA descriptor / TMA swizzle candidates on one small GEMM and print the
relative error of each, so a single gpurun call resolves the layout question."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infinite_video_b200 import _capi, ops

lib = _capi.lib()
lib.ltm_debug_set_mn_desc.argtypes = [C.c_uint] * 5
lib.ltm_debug_set_mn_desc.restype = None
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
cands = {
    "A layout1 lbo4096 sbo512  kadv1024 ATOM_32B": (1, 4096, 512, 1024, 4),
    "B layout1 lbo512  sbo4096 kadv1024 ATOM_32B": (1, 512, 4096, 1024, 4),
    "C layout1 lbo4096 sbo512  kadv1024 ATOM_32B_FLIP8": (1, 4096, 512, 1024, 5),
    "D layout2 lbo4096 sbo1024 kadv1024 SW128": (2, 4096, 1024, 1024, 3),
    "E layout1 lbo4096 sbo1024 kadv1024 ATOM_32B": (1, 4096, 1024, 1024, 4),
    "F layout1 lbo4096 sbo512  kadv1024 SW128": (1, 4096, 512, 1024, 3),
    "G layout2 lbo4096 sbo512  kadv1024 ATOM_32B": (2, 4096, 512, 1024, 4),
}
for bmn, amn in ((True, False), (False, True), (True, True)):
    M, N, K = 128, 128, 64
    A = torch.randn(*((K, M) if amn else (M, K)), generator=g)
    B = torch.randn(1, *((K, N) if bmn else (N, K)), generator=g)
    want = (A.t() if amn else A).double() @ (B[0] if bmn else B[0].t()).double()
    for name, c in cands.items():
        lib.ltm_debug_set_mn_desc(*c)
        try:
            got = ops.gemm(A.to(dev), B.to(dev), a_kmajor=not amn, b_kmajor=not bmn, precision="tf32x3")
            torch.cuda.synchronize()
            err = float((got[0].cpu().double() - want).abs().max() / want.abs().max())
        except Exception as e:  # noqa: BLE001
            err = f"EXC {e}"
        print(f"A_mn={amn} B_mn={bmn}  {name}: {err}")
lib.ltm_debug_set_mn_desc(1, 4096, 512, 1024, 4)
