#!/usr/bin/env python
"""GPU probe: accuracy / time of the short-term attention (N1) per precision choice."""
import os, sys, copy, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infinite_video_b200.cross_attention import CrossAttentionLTM
dev = torch.device("cuda:0")
torch.manual_seed(21)
lq, lk, lv = torch.nn.Linear(768, 768), torch.nn.Linear(768, 768), torch.nn.Linear(768, 768)
g = torch.Generator().manual_seed(22)
B, L = int(sys.argv[1]) if len(sys.argv) > 1 else 2, int(sys.argv[2]) if len(sys.argv) > 2 else 256
hidden = torch.randn(B, 32, 768, generator=g); enc = torch.randn(B, L * 32, 768, generator=g)
q = torch.nn.functional.linear(hidden, lq.weight, lq.bias)
def ref():
    H, d = 12, 64
    qh = q.double().view(B, 32, H, d).permute(0, 2, 1, 3)
    K = (enc.double() @ lk.weight.double().t() + lk.bias.double()).view(B, -1, H, d).permute(0, 2, 1, 3)
    V = (enc.double() @ lv.weight.double().t() + lv.bias.double()).view(B, -1, H, d).permute(0, 2, 1, 3)
    p = torch.softmax(qh @ K.transpose(-1, -2) / 8.0, -1)
    return (p @ V).permute(0, 2, 1, 3).reshape(B, 32, 768)
want = ref()
for sp, vp in (("tf32", "tf32"), ("tf32x3", "tf32"), ("tf32", "tf32x3"), ("tf32x3", "tf32x3")):
    m = CrossAttentionLTM(copy.deepcopy(lq).to(dev), copy.deepcopy(lk).to(dev), copy.deepcopy(lv).to(dev), 0.5, 256, .75,
                          score_precision=sp, value_precision=vp)
    got = m.short_term(q.to(dev), enc.to(dev)).cpu().double()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    qd, ed = q.to(dev), enc.to(dev)
    e0.record()
    for _ in range(5): m.short_term(qd, ed)
    e1.record(); torch.cuda.synchronize()
    print(sp, vp, "relerr", float((got - want).abs().max() / want.abs().max()), "ms per call (B=2)", e0.elapsed_time(e1) / 5)
