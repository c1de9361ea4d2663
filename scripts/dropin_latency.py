#!/usr/bin/env python
"""Wall-clock per call of the drop-in module used the way the reference pipeline uses it: ONE video, one forward per
chunk from Python (module overhead + uniform draws on the host + kernels), NExT-QA chunk shape."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infinite_video_b200 import LongTermAttention
dev = torch.device("cuda:0")
E, D, Q, L, T, N = 768, 768, 32, 256, 32, 256
torch.manual_seed(0)
key, val = torch.nn.Linear(E, D).to(dev), torch.nn.Linear(E, D).to(dev)
m = LongTermAttention(attn_num_basis=N, head_size=64, length=768, target_len=768, attn_func="softmax",
                      infinite_memory=True, n_layers=2, attn_drop=0.1, n_heads=12, d_model=768, affines=True, mask=True,
                      mask_type="cnn", kl_regularizer=False, sigma_0=None, mu_0=None, sticky_memories=True,
                      continuous=True, sigmas=None, tau=0.75, proj_key=key, proj_value=val)
ks = [torch.randn(1, L * T, E, device=dev) for _ in range(8)]
qs = [torch.randn(1, Q, D, device=dev) for _ in range(8)]
def video():
    for c in range(8):
        out = m(ks[c], qs[c], new_doc=(c == 0), layer_n=0)
    return out
for _ in range(3):
    video()
torch.cuda.synchronize()
t0 = time.perf_counter()
R = 20
for _ in range(R):
    video()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / (R * 8)
print(f"drop-in LongTermAttention.forward, 1 video, NExT-QA chunk: {dt * 1e6:.1f} us wall per call ({1 / dt:.0f} chunks/s)")
