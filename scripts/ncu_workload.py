"""Small workload that launches every kernel of the library once or twice, for `ncu -k regex:<kernel>` captures
(scripts/ncu_all.sh).  32 videos of the NExT-QA shape: rect path in tf32 (tensor-core attention) and tf32x3
(FMA attention), the Gaussian variant (ridge solve, RBF design, erf histogram, sorted re-sampling, Gaussian attention)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infinite_video_b200.batched import BatchedGaussLTM, BatchedRectLTM

dev = torch.device("cuda:0")
Bv, L, T, E, Q, N = int(os.environ.get("BV", 32)), 256, 32, 768, 32, 256
torch.manual_seed(0)
key, val = torch.nn.Linear(E, 768), torch.nn.Linear(E, 768)
w = (key.weight.detach(), key.bias.detach(), val.weight.detach(), val.bias.detach())
g = torch.Generator(device=dev).manual_seed(1)
k = [torch.randn(Bv, L * T, E, device=dev, generator=g) for _ in range(3)]
q = [torch.randn(Bv, Q, 768, device=dev, generator=g) for _ in range(3)]
u = [torch.rand(Bv, 512, device=dev, dtype=torch.float64, generator=g) for _ in range(3)]
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "rect"):
    for prec, kvd in (("tf32", "fp16"), ("tf32", "fp32"), ("tf32x3", "fp32")):
        eng = BatchedRectLTM(N, .75, *w, device=dev, precision=prec, keep_scores=(prec == "tf32x3"), kv_dtype=kvd)
        for c in range(3):
            eng.step(k[c], q[c], u[c] if c else None, new_doc=(c == 0))
        if prec == "tf32x3":
            eng.density()
    eng = BatchedRectLTM(512, .75, *w, device=dev, kv_dtype="fp32")            # num_basis 512 (fp32 K|V)
    for c in range(2):
        eng.step(k[c], q[c], u[c] if c else None, new_doc=(c == 0))
    eng = BatchedRectLTM(512, .75, *w, device=dev)            # num_basis 512
    for c in range(2):
        eng.step(k[c], q[c], u[c] if c else None, new_doc=(c == 0))
if which in ("all", "bins"):
    # per-bin pooling of update chunks (forced: 32 videos would not pick it by themselves)
    eng = BatchedRectLTM(N, .75, *w, device=dev, bin_pool=True)
    for c in range(3):
        eng.step(k[c], q[c], u[c] if c else None, new_doc=(c == 0))
if which in ("all", "caller"):
    # caller branch (N1): pooling + fp16 copy of the chunk in one pass, fp16 GEMMs, register-resident row softmax
    from infinite_video_b200.cross_attention import CrossAttentionLTM
    lin = lambda: torch.nn.Linear(E, 768).to(dev)
    mod = CrossAttentionLTM(lin(), lin(), lin(), alpha=0.9, num_basis=N, tau=.75, sticky=True, n_heads=12)
    for c in range(2):
        mod(q[c], k[c], new_video=(c == 0), u=u[c] if c else None)
if which in ("all", "gauss"):
    kg = [x.view(Bv, L, T, E)[:, :, 0].contiguous() for x in k]
    # default: folded sticky operator, fp16x2 projection, tensor-core attention; then the round-1 route (gathered
    # samples, split-TF32 projection, FMA attention) so that its kernels are captured too
    for kw in ({}, dict(fold_samples=False, proj_precision="tf32x3", tc_attn=False)):
        eng = BatchedGaussLTM(N, .75, *w, device=dev, **kw)
        for c in range(3):
            eng.step(kg[c], q[c], u[c] if c else None, new_doc=(c == 0))
torch.cuda.synchronize()
print("ok")
