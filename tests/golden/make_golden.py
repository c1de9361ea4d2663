"""Generates tests/golden/*.npz by running the UNMODIFIED reference modules from /root/reference
(dev container only).  Inputs are reproducible from the recorded seeds (torch CPU generator), so only
outputs are stored; wide tensors are stored as column slices to keep the fixtures small.

    python tests/golden/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader as RL  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
B_COLS = 96          # columns of B_past kept


def make_inputs(seed, C, Bv, rows, e, Q, q_scale=1.0):
    """Shared with the tests: chunk inputs from a private generator."""
    g = torch.Generator().manual_seed(seed)
    ks = [torch.randn(Bv, rows, e, generator=g) for _ in range(C)]
    qs = [torch.randn(Bv, Q, 768, generator=g) * q_scale for _ in range(C)]
    us = [torch.rand(Bv, 512, dtype=torch.float64, generator=g) for _ in range(C)]
    return ks, qs, us


def make_proj(seed, e):
    torch.manual_seed(seed)
    return nn.Linear(e, 768), nn.Linear(e, 768)


def run_gibbs(flavour, N, L, C, seed, tau=0.75, q_scale=1.0):
    T, e, Q = (32, 768, 32) if flavour == "vl" else (196, 1024, 96)
    mod = RL.load_gibbs_vl() if flavour == "vl" else RL.load_gibbs_vc()
    key, val = make_proj(seed, e)
    m = mod.LongTermAttention(**RL.caller_kwargs(N, tau, True, key, val))
    ks, qs, _ = make_inputs(seed + 1, C, 1, L * T, e, Q, q_scale)
    ctxs, Bs, bs, ps, us = [], [], [], [], []
    # the sampled bins are a local of update_inf (gibbs:204-205): observe them through Categorical.sample without
    # touching the reference source.  Per sticky call the first sample() is `b` [512,B], the second the dummy jitter.
    rec = []
    orig_sample = torch.distributions.Categorical.sample

    def spy(self, sample_shape=torch.Size()):
        out = orig_sample(self, sample_shape)
        rec.append((self.probs.detach().clone(), out.detach().clone()))
        return out
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())          # the VL copy pickles ./alphas_uniform on every call
    try:
        with torch.no_grad():
            for c in range(C):
                torch.manual_seed(5000 + seed + c)
                rec.clear()
                torch.distributions.Categorical.sample = spy
                try:
                    ctx = m(ks[c], qs[c], new_doc=(c == 0), layer_n=0)
                finally:
                    torch.distributions.Categorical.sample = orig_sample
                if c == 0:
                    assert not rec
                    bs.append(np.full((1, 512), -1, dtype=np.int64))
                    ps.append(np.zeros((1, 127), dtype=np.float32))
                else:
                    assert len(rec) == 2 and rec[0][1].shape == (512, 1)
                    bs.append(rec[0][1].t().numpy().copy())                  # [1,512] bins in draw order
                    ps.append(rec[0][0].numpy().copy())                      # [1,127] probabilities sampled from
                torch.manual_seed(5000 + seed + c)
                us.append(torch.rand(1, 512, dtype=torch.float64).numpy())   # the draws the reference consumed
                ctxs.append(ctx[0].numpy().copy())
                Bs.append(m.B_past[0, :, :B_COLS].numpy().copy())
    finally:
        os.chdir(cwd)
    return dict(ctx=np.stack(ctxs), B_cols=np.stack(Bs), u=np.stack(us), b=np.stack(bs), p=np.stack(ps),
                meta=np.array([N, L, C, seed, T, e, Q], dtype=np.int64), tau=np.float64(tau),
                q_scale=np.float64(q_scale), B_absmean=np.array([float(np.abs(b).mean()) for b in Bs]))


def run_gauss(N, L, C, seed, Bv=1, tau=0.75):
    e, Q = 768, 32
    mod = RL.load_gaussian_vl()
    key, val = make_proj(seed, e)
    m = mod.LongTermAttention(**RL.caller_kwargs(N, tau, True, key, val, sigmas=[0.005, 0.01]))
    m.device = "cpu"
    ks, qs, _ = make_inputs(seed + 1, C, Bv, L, e, Q)
    ctxs, Bs, us, bs = [], [], [], []
    rec = []
    orig_sample = torch.distributions.Categorical.sample

    def spy(self, sample_shape=torch.Size()):
        out = orig_sample(self, sample_shape)
        rec.append(out.detach().clone())
        return out
    with torch.no_grad():
        for c in range(C):
            m.length = m.target_len = L
            torch.manual_seed(7000 + seed + c)
            rec.clear()
            torch.distributions.Categorical.sample = spy
            try:
                ctx = m(ks[c], qs[c], new_doc=(c == 0), layer_n=0)
            finally:
                torch.distributions.Categorical.sample = orig_sample
            bs.append(rec[0].t().numpy().copy() if c else np.full((Bv, 512), -1, dtype=np.int64))   # draw order
            torch.manual_seed(7000 + seed + c)
            nn.Linear(N, 1, bias=False); nn.Linear(N, 1, bias=False)      # replay the throw-away inits (:92-95)
            us.append(torch.rand(Bv, 512, dtype=torch.float64).numpy())
            ctxs.append(ctx.numpy().copy())
            Bs.append(m.B_past[:, :, :B_COLS].numpy().copy())
    return dict(ctx=np.stack(ctxs), B_cols=np.stack(Bs), u=np.stack(us), b=np.stack(bs), G_inf=m.G_inf.numpy().copy(),
                G0=m.Gs[L].numpy().copy(), meta=np.array([N, L, C, seed, Bv, e, Q], dtype=np.int64),
                tau=np.float64(tau))


if __name__ == "__main__":
    torch.set_num_threads(8)
    np.savez_compressed(os.path.join(OUT, "gibbs_vl_cfg1.npz"), **run_gibbs("vl", 64, 8, 4, seed=11))
    np.savez_compressed(os.path.join(OUT, "gibbs_vl_cfg2.npz"), **run_gibbs("vl", 256, 256, 3, seed=12))
    np.savez_compressed(os.path.join(OUT, "gibbs_vl_peaky.npz"), **run_gibbs("vl", 64, 8, 3, seed=13, q_scale=8.0))
    np.savez_compressed(os.path.join(OUT, "gibbs_vc_cfg3.npz"), **run_gibbs("vc", 64, 16, 2, seed=14))
    np.savez_compressed(os.path.join(OUT, "gibbs_vl_cfg4.npz"), **run_gibbs("vl", 512, 256, 3, seed=15))
    np.savez_compressed(os.path.join(OUT, "gauss_small.npz"), **run_gauss(64, 8, 3, seed=21, Bv=2))
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
