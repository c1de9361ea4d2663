import os, sys
sys.path.insert(0, "/root/repo")
import torch
from infinite_video_b200 import ops, tables
dev = torch.device("cuda:0")
torch.manual_seed(0)
rel = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()
for (Bv, N, Q, H, L) in ((3, 256, 1, 12, 64), (2, 128, 96, 16, 16), (1, 256, 33, 8, 32), (5, 128, 32, 1, 16)):
    D = H * 64
    tab = tables.rect_tables(L, N, .75); td = tab.to(dev)
    KV = torch.randn(Bv, N, 2 * D, device=dev)
    KVr = ((KV.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    q = torch.randn(Bv, Q, D, device=dev) * 2
    Kt = KVr[:, :, :D].reshape(Bv, N, H, 64).permute(0, 2, 3, 1).contiguous(); V = KVr[:, :, D:].contiguous()
    c0, s0, h0 = ops.cont_attn_rect_t(q, Kt, V, td["W"], tab.W_out, td["jb"], td["tb"], want_scores=True)
    c1, s1, h1 = ops.cont_attn_rect_tc(q, KVr, td["X"], td["W"], tab.W_out, tab.c_none, td["jb"], td["tb"], want_scores=True, n_heads=H)
    torch.cuda.synchronize()
    print((Bv, N, Q, H, L), "scores", f"{rel(s1, s0):.2e}", "ctx", f"{rel(c1, c0):.2e}", "hist", f"{rel(h1, h0):.2e}", "finite", bool(torch.isfinite(c1).all()))
