#!/usr/bin/env python
"""One K/V-projection GEMM launch pattern for ncu (-k regex:gemm_tf32 -s 3 -c 1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infinite_video_b200 import ops
dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
A = torch.randn(32768, 768, device=dev); W = torch.randn(1536, 768, device=dev); bias = torch.zeros(1536, device=dev)
out = torch.empty(32768, 1536, device=dev)
for _ in range(5):
    ops.project_kv(A, W, bias, prec, out=out)
torch.cuda.synchronize()
