"""N1: the cross-attention branch of the Q-former's `BertSelfAttention` -- short-term softmax
attention over the chunk's L*T encoder tokens plus the (1 - alpha) LTM context -- on libinfltm.

Reference: infty-Video-LLaMA/InfVideoLLaMA/models/Qformer.py:197-310
    q = query(hidden)                                   :211
    a_long = LTM(enc, q, new_video)   iff alpha != 1    :216-223
    K, V = key(enc), value(enc)                         :225-226     (two [L*T, e] x [e, 768] GEMMs)
    probs = softmax(q_h K_h^T / sqrt(d) + mask)         :241, :279-285
    ctx = probs V                                       :297-300
    out = alpha * ctx + (1 - alpha) * a_long            :303-304

The short-term half is re-associated so that K and V of the 8192 tokens are never materialised:
    scores_h = (q_h W_k,h) enc^T / sqrt(d)      (+ q_h.b_k,h, a per-row constant: softmax-invariant, dropped)
    ctx_h    = (probs_h enc) W_v,h^T + b_v,h    (rows of probs sum to 1)
which halves the flops (9.7 instead of 19.3+ GFLOP per video-chunk at L*T = 8192) and removes 100 MB of K/V
traffic per video.  Every contraction runs on the tcgen05 GEMM (`ltm_gemm`), the row softmax and the blend are two
small streaming kernels.  The key/value projections are the ones the LTM shares (`proj_key=self.key`, :156-157).
"""
import math

import torch
import torch.nn as nn

from . import ops
from .ltm import LongTermAttention


class CrossAttentionLTM(nn.Module):
    """Drop-in for the cross-attention use of `BertSelfAttention` (eval mode, absolute position embeddings, no head
    mask).  Construct from the caller's own `query` / `key` / `value` `nn.Linear`s and its config values."""

    def __init__(self, query: nn.Linear, key: nn.Linear, value: nn.Linear, alpha: float, num_basis: int, tau: float,
                 sticky: bool = True, n_heads: int = 12, tokens_per_frame: int = 32, videos_per_pass: int = 32,
                 precision: str = "tf32", score_precision: str = None, value_precision: str = None,
                 operands: str = "fp16", **ltm_kwargs):
        super().__init__()
        self.query, self.key, self.value = query, key, value
        self.alpha = float(alpha)
        self.H = n_heads
        self.d = query.out_features // n_heads
        # bounds the [videos, H*Q, L*T] score buffer (12.6 MB per video at L*T = 8192); large passes amortise the
        # host side of the five GEMM launches, which is what bounds small batches
        self.videos_per_pass = videos_per_pass
        self.precision = precision
        # the two large contractions (scores = Qt enc^T, Y = probs enc); the three small ones always run split-TF32.
        # Measured (scripts/stm_probe.py, error vs fp64 at L*T = 8192 / 256, time at 16 videos):
        #   scores x1 (Qt rounded to tf32) / values x3   2.6e-4 / 2.1e-4   0.80 ms   <- default
        #   scores x3 / values x1                         4.5e-4 / 9.4e-4   0.87 ms
        #   x1 / x1                                       5.8e-4 / 9.2e-4   0.59 ms
        #   x3 / x3                                       3.9e-5 / 1.2e-5   1.08 ms
        # (the value contraction reads the raw chunk, which the tensor core truncates: its bias is the larger error)
        self.score_precision = score_precision or precision
        self.value_precision = value_precision or "tf32x3"
        # operands of the two large contractions.  "fp16" (default): the chunk tokens are converted ONCE per chunk to
        # fp16 with round-to-nearest (one streaming pass, 25 -> 12.6 MB per video), `Qt` and the probabilities are
        # produced as fp16, and both contractions run as kind::f16 UMMAs (twice the TF32 rate, half the operand bytes;
        # the values need no split: rounding is unbiased where the tensor core's truncation of fp32 was not).
        # "fp32": the TF32 / split-TF32 path described above.
        if operands not in ("fp16", "fp32"):
            raise ValueError("operands must be 'fp16' or 'fp32'")
        self.operands = operands
        self.long_term_attention = LongTermAttention(
            head_size=self.d, length=key.in_features, target_len=key.in_features, attn_func="softmax",
            attn_num_basis=num_basis, continuous=True, attn_drop=0.1, infinite_memory=True, n_layers=2,
            n_heads=n_heads, affines=True, mask=True, mask_type="cnn", kl_regularizer=False, proj_key=key,
            proj_value=value, sigma_0=None, mu_0=None, sticky_memories=sticky, sigmas=None, tau=tau,
            d_model=query.out_features, tokens_per_frame=tokens_per_frame, **ltm_kwargs)

    def _linear(self, x, lin, precision):
        """x[M, in] @ W^T + b on tcgen05."""
        M, K = x.shape
        out = torch.empty(M, lin.out_features, device=x.device, dtype=torch.float32)
        w = lin.weight.detach().float().contiguous()
        b = None if lin.bias is None else lin.bias.detach().float().contiguous()
        return ops.gemm_raw(x, K, 0, True, w, K, 0, True, out, lin.out_features, 0, M, lin.out_features, K, 1,
                            bias=b, precision=precision)

    @torch.no_grad()
    def short_term(self, q, enc, mask=None, enc16=None):
        """softmax(q_h K_h^T / sqrt(d) + mask) V_h for all heads.  q[B,Q,D] (already projected), enc[B,LT,e] fp32,
        mask: additive [B, LT] or None; enc16: the chunk as float16 when the caller already has it (the pooling pass of
        the LTM writes it, `LongTermAttention.pool_for_caller`).  Returns [B,Q,D]."""
        B, Q, D = q.shape
        LT, e = enc.shape[1], enc.shape[2]
        H, d = self.H, self.d
        dev = q.device
        wk = self.key.weight.detach().float().contiguous()          # [D, e]
        wv = self.value.weight.detach().float().contiguous()
        bv = self.value.bias.detach().float().contiguous() if self.value.bias is not None else None
        out = torch.empty(B, Q, D, device=dev, dtype=torch.float32)
        for v0 in range(0, B, self.videos_per_pass):
            nb = min(self.videos_per_pass, B - v0)
            qv = q[v0:v0 + nb].contiguous()
            ev = enc[v0:v0 + nb].contiguous()
            if self.operands == "fp16" and e % 8 == 0 and LT % 8 == 0:
                # [nb, LT, e] fp16, rounded: from the LTM's pooling pass, or converted here
                e16 = enc16[v0:v0 + nb].contiguous() if enc16 is not None else ops.to_half(ev)
                # (1) Qt as fp16, straight out of the (split-TF32) GEMM epilogue
                Qt = torch.empty(nb, H, Q, e, device=dev, dtype=torch.float16)
                ops.gemm_raw(qv, D, d, True, wk, e, d * e, False, Qt, e, Q * e, nb * Q, e, d, H,
                             c_group=Q, c_group_stride=H * Q * e, precision="tf32x3", c_fp16=True)
                # (2) scores[v] = Qt[v] enc[v]^T : both operands fp16, K-major
                S = torch.empty(nb, H * Q, LT, device=dev, dtype=torch.float32)
                ops.gemm_raw(Qt, e, H * Q * e, True, e16, e, LT * e, True, S, LT, H * Q * LT, H * Q, LT, e, nb,
                             precision="tf32", ab_fp16=True)
                # (3) probabilities as fp16
                m = None if mask is None else mask[v0:v0 + nb].float().contiguous()
                P = ops.softmax_rows_half(S, 1.0 / math.sqrt(d), m, H * Q)
                # (4) Yh[h][v][q][:] = probs enc[v] : A fp16 K-major, B = enc16[v] read as [K = LT][N = e] (MN-major)
                Yh = torch.empty(H, nb, Q, e, device=dev, dtype=torch.float32)
                ops.gemm_raw(P, LT, H * Q * LT, True, e16, e, LT * e, False, Yh, e, Q * e, H * Q, e, LT, nb,
                             c_group=Q, c_group_stride=nb * Q * e, precision="tf32", ab_fp16=True)
                # (5) as below
                ops.gemm_raw(Yh, e, nb * Q * e, True, wv, e, d * e, True, out, D, d, nb * Q, d, e, H,
                             bias=bv, bias_stride=d, precision="tf32x3", c_offset=v0 * Q * D)
                continue
            # (1) Qt[v][h][q][:] = q_h W_k,h : batch over heads, A = column block h of q, B = row block h of W_k
            #     read as [K = d][N = e] (MN-major); rows m = (v, q) land at (v*H + h)*Q + q
            Qt = torch.empty(nb, H, Q, e, device=dev, dtype=torch.float32)
            ops.gemm_raw(qv, D, d, True, wk, e, d * e, False, Qt, e, Q * e, nb * Q, e, d, H,
                         c_group=Q, c_group_stride=H * Q * e, precision="tf32x3",
                         round_tf32=(self.score_precision == "tf32"))   # rounded (not truncated) operand of (2)
            # (2) scores[v] = Qt[v] enc[v]^T : [H*Q, e] x [e, LT]
            S = torch.empty(nb, H * Q, LT, device=dev, dtype=torch.float32)
            ops.gemm_raw(Qt, e, H * Q * e, True, ev, e, LT * e, True, S, LT, H * Q * LT, H * Q, LT, e, nb,
                         precision=self.score_precision)
            # (3) probs = softmax(scores / sqrt(d) + mask), in place
            m = None if mask is None else mask[v0:v0 + nb].float().contiguous()
            ops.softmax_rows(S, 1.0 / math.sqrt(d), m, H * Q)
            # (4) Yh[h][v][q][:] = probs[v][(h,q)] enc[v] : B operand = enc[v] read as [K = LT][N = e] (MN-major)
            Yh = torch.empty(H, nb, Q, e, device=dev, dtype=torch.float32)
            ops.gemm_raw(S, LT, H * Q * LT, True, ev, e, LT * e, False, Yh, e, Q * e, H * Q, e, LT, nb,
                         c_group=Q, c_group_stride=nb * Q * e, precision=self.value_precision)
            # (5) ctx[:, h*d:(h+1)*d] = Yh[h] W_v,h^T + b_v,h : batch over heads
            ops.gemm_raw(Yh, e, nb * Q * e, True, wv, e, d * e, True, out, D, d, nb * Q, d, e, H,
                         bias=bv, bias_stride=d, precision="tf32x3", c_offset=v0 * Q * D)
        return out

    @torch.no_grad()
    def forward(self, hidden_states, encoder_hidden_states, new_video=False, layer=None,
                encoder_attention_mask=None, u=None):
        """Returns the blended context layer [B,Q,D] (what `BertSelfAttention.forward` returns as outputs[0])."""
        if not hidden_states.is_cuda:
            raise RuntimeError("CrossAttentionLTM runs on CUDA tensors only (no CPU fallback)")
        B, Q, _ = hidden_states.shape
        enc = encoder_hidden_states.float().contiguous()
        q = self._linear(hidden_states.float().reshape(B * Q, -1).contiguous(), self.query, "tf32x3")
        q = q.view(B, Q, -1)
        mask = None
        if encoder_attention_mask is not None:                       # HF additive mask [B,1,1,LT] -> [B,LT]
            mask = encoder_attention_mask.reshape(B, -1)
        ltm = self.long_term_attention
        enc16 = None
        if (self.alpha != 1.0 and self.operands == "fp16" and ltm.variant == "gibbs" and ltm.share_pooling
                and enc.dtype == torch.float32 and enc.is_contiguous() and enc.shape[2] % 8 == 0 and enc.shape[1] % 8 == 0
                and enc.shape[1] % ltm.tokens_per_frame == 0):
            # one pass over the chunk: pooled frames for the LTM (parked in its shared-pooling slot) + the fp16 copy
            # the short-term GEMMs read
            enc16 = ltm.pool_for_caller(enc)
        ctx = self.short_term(q, enc, mask, enc16=enc16)
        if self.alpha == 1.0:
            return ctx.to(hidden_states.dtype)
        a_long = self.long_term_attention(enc, q, new_doc=new_video, layer_n=layer, u=u)
        return ops.blend(ctx, a_long.float(), self.alpha).to(hidden_states.dtype)
