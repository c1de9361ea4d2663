"""Host constant tables (product code) against the oracle's dense reference operators."""
import numpy as np
import pytest
import torch

from infinite_video_b200 import tables as T
from oracle import ltm_oracle as O
from tests.helpers import make_inputs, make_proj, proj_tensors

SHAPES = [(64, 8, .75), (256, 256, .75), (512, 256, .75), (64, 16, .75), (256, 256, .9), (100, 30, .5),
          (64, 7, .75), (64, 2, .75), (128, 1024, .75), (64, 3, .6)]


@pytest.mark.parametrize("N,L,tau", SHAPES)
def test_segmented_mean_tables_equal_dense_ridge_operators(N, L, tau):
    t = T.rect_tables(L, N, tau)
    psi = O.RectBasis(N)
    G0 = O.ridge_operator(psi, O.first_chunk_positions(L), L).numpy()
    pos, _ = O.update_positions(L, tau)
    Gi = O.ridge_operator(psi, pos, 512 + L).numpy()
    assert np.array_equal(t.dense_G0(), G0)           # bitwise: 1/(cnt+0.5) in fp32
    assert np.array_equal(t.dense_Ginf(), Gi)
    assert (Gi[-1] == 0).all()                        # the last new frame (position 1.0) is dropped
    assert t.idx_uniform[-1] == -1                    # uniform table: t/tau == 1.0 hits no basis
    assert t.jb[0] == -1 and t.jb[-1] == -1           # nudged end edges hit no basis


@pytest.mark.parametrize("N", [64, 128, 256, 512])
def test_sticky_positions_map_to_floor_bins_for_pow2(N):
    t = T.rect_tables(8, N, .75)
    want = np.floor(np.arange(128) / 128.0 * N).astype(np.int32)
    assert np.array_equal(t.bin2basis, want)


@pytest.mark.parametrize("N,L", [(64, 8), (256, 32), (512, 16), (100, 30)])
def test_closed_form_quadrature_and_histogram(N, L):
    """r_j = W_j e^{S_j}/(sum W e^S + W_out) == the reference's 1000-point trapezoid integral; the closed-form
    sticky histogram == cumulative_trapezoid/diff of the reference."""
    key, val = make_proj(1, 768)
    orc = O.RectLTM(N, .75, *proj_tensors(key, val))
    ks, qs, us = make_inputs(2, 2, 1, L * 32, 768, 32, q_scale=4.0)
    t = T.rect_tables(L, N, .75)
    with torch.no_grad():
        orc.forward(ks[0], qs[0], True)
        S = orc.last["scores"].double()
        W = torch.from_numpy(t.W).double()
        num = W * torch.exp(S)
        r = num / (num.sum(-1, keepdim=True) + t.W_out)
        err = float((r - orc.last["r"].double()).abs().max() / orc.last["r"].abs().max())
        assert err < 2e-5, err
        # histogram
        p_ref = orc.sticky_hist(orc.tables(L))[0].double()
        jb = torch.from_numpy(t.jb).long()
        tb = torch.from_numpy(t.tb).double()
        z = torch.where(jb >= 0, S[..., jb.clamp(min=0)], torch.zeros((), dtype=torch.float64))
        E = torch.exp(z)
        dt = tb[1:] - tb[:-1]
        Z = (dt * (E[..., 1:] + E[..., :-1]) / 2).sum(-1, keepdim=True)
        inc = dt * (E[..., 1:] + E[..., :-1]) / 2 / Z
        p = inc[..., 1:].sum((1, 2))[0]
        p = p / p.sum()
        assert float((p - p_ref).abs().max() / p_ref.max()) < 2e-5


def test_single_frame_chunks_are_rejected():
    with pytest.raises(ValueError):
        T.rect_tables(1, 64, .75)


def test_gauss_tables_shapes():
    t = T.gauss_tables(8, 63, .75)
    assert t.N == 64 and t.pos0.shape[0] == 16 and t.pos1.shape[0] == 2 * (512 + 8)
    assert t.trim1 == 260 and t.trim0 == 4
    with pytest.raises(ValueError):
        T.gauss_tables(7, 64, .75)


@pytest.mark.parametrize("N,L", [(256, 256), (128, 16)])
def test_tensor_core_attention_operand_rows(N, L):
    """tables.X (rows appended to V^T by csrc/attn_tc.cu): with e_j = W_j exp(S_j - m) as the other operand, column 0
    gives the quadrature normaliser and columns 1 + 2 (hi + lo, hi exact in tf32) the trapezoid integral of the
    sticky-edge density, Z = trapz(E, tb) with E_i = exp(S[jb_i] - m) (0 score where no basis is active)."""
    t = T.rect_tables(L, N, .75)
    assert t.X is not None and t.X.shape == (N, 32) and (t.X[:, 0] == 1).all() and (t.X[:, 3:] == 0).all()
    hi_bits = t.X[:, 1].view(np.uint32)
    assert (hi_bits & np.uint32(0x1FFF) == 0).all()                    # exact on the tf32 grid
    rng = np.random.default_rng(3)
    S = rng.normal(size=N) * 3.0
    m = max(0.0, S.max())
    E = np.where(t.jb >= 0, np.exp(S[np.maximum(t.jb, 0)] - m), np.exp(-m))
    tb = t.tb.astype(np.float64)
    want = float(np.sum((tb[1:] - tb[:-1]) * (E[1:] + E[:-1]) * 0.5))
    e = t.W.astype(np.float64) * np.exp(S - m)
    got = float(((t.X[:, 1].astype(np.float64) + t.X[:, 2]) * e).sum() + t.c_none * np.exp(-m))
    assert abs(got - want) < 1e-6 * want
    assert abs(float((t.X[:, 0] * e).sum()) - float(e.sum())) < 1e-12


def test_log_spacing_tables_equal_the_dense_operator():
    """spacing='log' moves the first-chunk frame positions (gibbs:114-127): CSR tables == dense ridge operator."""
    for N, L in ((64, 8), (256, 32), (64, 7), (100, 30)):
        t = T.rect_tables(L, N, .75, spacing="log")
        psi = O.RectBasis(N)
        G0 = O.ridge_operator(psi, O.first_chunk_positions(L, "log"), L).numpy()
        assert np.array_equal(t.dense_G0(), G0), (N, L)
        assert not np.array_equal(t.dense_G0(), T.rect_tables(L, N, .75).dense_G0())


@pytest.mark.parametrize("N,L,spacing", [(256, 256, "linear"), (512, 256, "linear"), (64, 8, "linear"), (64, 16, "log"),
                                         (100, 64, "linear"), (128, 256, "log"), (64, 5, "linear")])
def test_folded_frame_tables_expand_to_the_per_frame_tables(N, L, spacing):
    """`fbin_ptr` / `seg_mem1b` (per-bin pooling, csrc/pool.cu) are the update tables with the frames of a bin collapsed
    into one member: expanding them must give back `seg_mem1` member for member, in order."""
    t = T.rect_tables(L, N, 0.75, 512, spacing=spacing)
    assert t.xb_rows == N - t.jf and t.xb_row0 == t.jf
    assert t.fbin_ptr.shape[0] == t.xb_rows + 1 and (np.diff(t.fbin_ptr) >= 0).all()
    for j in range(N):
        want = list(t.seg_mem1[t.seg_ptr1[j]:t.seg_ptr1[j + 1]])
        got = []
        for m in t.seg_mem1b[t.seg_ptr1b[j]:t.seg_ptr1b[j + 1]]:
            if m < t.S:
                got.append(int(m))
            else:
                r = int(m) - t.S
                assert r == j - t.xb_row0
                got.extend(t.S + f for f in range(t.fbin_ptr[r], t.fbin_ptr[r + 1]))
        assert got == want, j


def test_folded_frame_tables_over_many_shapes():
    """Property over a spread of (L, N, tau, spacing): either the frames cannot be folded (xb_rows == 0: the engine keeps
    the per-frame layout) or the folded tables expand to the per-frame tables exactly and cover every frame that lies
    inside a basis once."""
    rng = np.random.default_rng(5)
    seen = 0
    for _ in range(60):
        L = int(rng.integers(2, 300))
        N = int(rng.choice([16, 32, 64, 100, 128, 256, 512]))
        tau = float(rng.choice([0.5, 0.6, 0.75, 0.9]))
        spacing = str(rng.choice(["linear", "log"]))
        t = T.rect_tables(L, N, tau, 512, spacing=spacing)
        if t.xb_rows == 0:
            continue
        seen += 1
        assert t.xb_row0 == t.jf and t.xb_rows == N - t.jf
        covered = np.zeros(L, bool)
        for r in range(t.xb_rows):
            assert not covered[t.fbin_ptr[r]:t.fbin_ptr[r + 1]].any()
            covered[t.fbin_ptr[r]:t.fbin_ptr[r + 1]] = True
        for j in range(N):
            want = list(t.seg_mem1[t.seg_ptr1[j]:t.seg_ptr1[j + 1]])
            got = []
            for m in t.seg_mem1b[t.seg_ptr1b[j]:t.seg_ptr1b[j + 1]]:
                if m < t.S:
                    got.append(int(m))
                else:
                    r = int(m) - t.S
                    got.extend(t.S + f for f in range(t.fbin_ptr[r], t.fbin_ptr[r + 1]))
            assert got == want, (L, N, tau, spacing, j)
        frames_in_a_bin = np.zeros(L, bool)
        for j in range(N):
            mem = t.seg_mem1[t.seg_ptr1[j]:t.seg_ptr1[j + 1]]
            frames_in_a_bin[mem[mem >= t.S] - t.S] = True
        assert np.array_equal(covered, frames_in_a_bin), (L, N, tau, spacing)
    assert seen >= 30
