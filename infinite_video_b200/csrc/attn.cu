// R10/R11 (+R6) and G4 -- continuous attention over the reconstructed signal, fused into one pass.
//
// Variant R (long_term_attention_gibbs.py:224-286): the reference evaluates the piecewise-constant
// score z(t) = S[bin(t)] on a 1000-point grid, normalises exp(z) with a trapezoid rule and integrates
// p(t) psi_j(t) numerically, materialising a [B,h,q,N,1000] integrand (393 MB at N=256).  Because z
// is constant inside a bin the quadrature collapses exactly to
//      r_j = W_j e^{S_j} / (sum_i W_i e^{S_i} + W_out)
// with W_j the summed trapezoid weights of the grid points inside bin j and W_out the weight of the
// grid points in no bin (t = 1.0), both constant tables (tables.py).  ctx = r V.  The same CTA also
// emits the sticky-histogram partials the *next* call needs (R6), so scores never travel to HBM.
//
// Variant G (long_term_attention.py:286-325): a = softmax(20 S); mu = a.mu_b;
// var = a.(mu_b^2 + sigma_b^2) - mu^2; r_j = N(mu; mu_j, sigma_j^2 + var)  (closed-form Gaussian x RBF
// integral, basis_functions.py:154-156,209-211); ctx = r V; (mu, sqrt(var)) saved for the next call.
//
// One CTA per (32-query tile, head, video).  K_h / V_h tiles of 64 basis rows stream through a two-slot
// cp.async ring (coalesced 16-byte copies of 256-byte row segments, the next tile in flight while the
// current one is consumed).  Each lane owns one query row: S = q K^T with the query row in registers and
// 4 K rows warp-broadcast per step; ctx = r V with all 64 output columns in registers and the basis
// dimension split across the 8 warps, reduced once through shared memory.
#include "common.cuh"
#include "rect_hist.cuh"

namespace ltm {

constexpr int DH = 64;        // head size (both Q-formers: 768 / 12)
constexpr int QT = 32;        // query rows per CTA
constexpr int JT = 64;        // basis rows per K/V tile
constexpr int ATTN_THREADS = 256;

struct AttnParams {
  const float* q;        // [Bv,Q,D]
  const float* KV;       // [Bv,N,2D]
  const float* tabA;     // rect: W[N]        gauss: basis_mu[N]
  const float* tabB;     // rect: unused      gauss: basis_sigma[N]
  float W_out;
  const int32_t* jb;     // rect hist
  const float* tb;
  float* ctx;            // [Bv,Q,D]
  float* scores_out;     // optional [Bv,H,Q,N]
  float* hist_part;      // rect: optional [Bv, H*q_tiles, 127]
  float* mu_out;         // gauss: [Bv,H*Q]
  float* sd_out;
  int Q, N, H;
};

static_assert(2 * JT * DH >= QT * (EDGES + 1), "histogram scratch must fit the (idle) tile ring");

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

// Shared-memory carve-up (floats).  The [8 warps][16][32] float4 reduction scratch of phase 3 aliases the
// score tile + the K/V tile ring once both are dead.
__host__ __device__ inline int attn_front_floats(int N) {
  const int ss = (QT * (N + 1) + 3) & ~3;
  const int need = ss + 2 * JT * DH;
  const int red = 8 * QT * DH;
  return need > red ? need : red;
}

template <int MODE>  // 0 = rect, 1 = gauss
__global__ void __launch_bounds__(ATTN_THREADS, 2)
cont_attn_kernel(const AttnParams p) {
  extern __shared__ __align__(16) float smem[];
  const int N = p.N, Q = p.Q, H = p.H, D = H * DH;
  const int SS = N + 1;                                  // padded row stride of the score tile
  float* Ss = smem;                                      // [QT][N+1]
  float* tiles = Ss + ((QT * SS + 3) & ~3);              // [2][JT][DH] cp.async ring (buffer 0 doubles as hist scratch)
  float* qs = smem + attn_front_floats(N);               // [QT][DH+1]
  float* tabA = qs + QT * (DH + 1);                      // [N]
  float* tabB = tabA + N;                                // [N]
  float* mrow = tabB + N;                                // [QT]
  float* zrow = mrow + QT;                               // [QT]
  float* part = zrow + QT;                               // [128]

  const int qt = blockIdx.x, h = blockIdx.y, v = blockIdx.z;
  const int q0 = qt * QT;
  const int rows = min(QT, Q - q0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* KVv = p.KV + (size_t)v * N * 2 * D;
  const int ntile = (N + JT - 1) / JT;

  // tile t in [0, 2*ntile): first the K_h tiles, then the V_h tiles; rows beyond N are zero-filled
  auto issue_tile = [&](int t) {
    const int isv = t >= ntile;
    const int j0 = (isv ? t - ntile : t) * JT;
    float* dst = tiles + (t & 1) * (JT * DH);
    const float* src = KVv + (isv ? D : 0) + h * DH;
    for (int f = tid; f < JT * (DH / 4); f += ATTN_THREADS) {
      const int r = f / (DH / 4), c4 = f - r * (DH / 4);
      if (j0 + r < N) cp_async16(dst + r * DH + 4 * c4, src + (size_t)(j0 + r) * 2 * D + 4 * c4);
      else reinterpret_cast<float4*>(dst + r * DH)[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
  };
  issue_tile(0);

  // ---- stage the query tile (scaled by 1/sqrt(d), gibbs:226) and the per-basis tables
  const float inv_sqrt_d = 1.0f / sqrtf((float)DH);      // d = 64 -> exactly 1/8
  for (int f = tid; f < QT * DH; f += ATTN_THREADS) {
    const int r = f / DH, c = f - r * DH;
    float val = 0.f;
    if (r < rows) val = p.q[((size_t)v * Q + q0 + r) * D + h * DH + c] * inv_sqrt_d;
    qs[r * (DH + 1) + c] = val;
  }
  for (int j = tid; j < N; j += ATTN_THREADS) {
    tabA[j] = p.tabA[j];
    if (MODE == 1) tabB[j] = p.tabB[j];
  }
  __syncthreads();

  // ---- phase 1: S[q, j] = q_h . K_h[j]; lane = query row (held in registers), each warp owns 8 of the 64
  //      tile rows as two groups of 4, K rows are warp-broadcast from shared memory
  {
    float qreg[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) qreg[c] = qs[lane * (DH + 1) + c];
    for (int t = 0; t < ntile; ++t) {
      if (t + 1 < ntile) { issue_tile(t + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
      __syncthreads();
      const float* tile = tiles + (t & 1) * (JT * DH);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int jj = warp * 8 + g * 4;
        const float4* k0 = reinterpret_cast<const float4*>(tile + (jj + 0) * DH);
        const float4* k1 = reinterpret_cast<const float4*>(tile + (jj + 1) * DH);
        const float4* k2 = reinterpret_cast<const float4*>(tile + (jj + 2) * DH);
        const float4* k3 = reinterpret_cast<const float4*>(tile + (jj + 3) * DH);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < DH / 4; ++c4) {
          const float4 x0 = k0[c4], x1 = k1[c4], x2 = k2[c4], x3 = k3[c4];
          const float qa = qreg[4 * c4], qb = qreg[4 * c4 + 1], qc = qreg[4 * c4 + 2], qd = qreg[4 * c4 + 3];
          a0 = fmaf(qa, x0.x, a0); a1 = fmaf(qa, x1.x, a1); a2 = fmaf(qa, x2.x, a2); a3 = fmaf(qa, x3.x, a3);
          a0 = fmaf(qb, x0.y, a0); a1 = fmaf(qb, x1.y, a1); a2 = fmaf(qb, x2.y, a2); a3 = fmaf(qb, x3.y, a3);
          a0 = fmaf(qc, x0.z, a0); a1 = fmaf(qc, x1.z, a1); a2 = fmaf(qc, x2.z, a2); a3 = fmaf(qc, x3.z, a3);
          a0 = fmaf(qd, x0.w, a0); a1 = fmaf(qd, x1.w, a1); a2 = fmaf(qd, x2.w, a2); a3 = fmaf(qd, x3.w, a3);
        }
        const int j = t * JT + jj;
        float* dst = Ss + lane * SS + j;
        if (j + 0 < N) dst[0] = a0;
        if (j + 1 < N) dst[1] = a1;
        if (j + 2 < N) dst[2] = a2;
        if (j + 3 < N) dst[3] = a3;
      }
      __syncthreads();                                   // tile (t&1) may be overwritten by the next prefetch
    }
  }
  // here: all scores are in Ss and the tile ring is idle (it serves as histogram scratch below)

  if (p.scores_out) {
    for (int f = tid; f < rows * N; f += ATTN_THREADS) {
      const int r = f / N, j = f - r * N;
      p.scores_out[(((size_t)v * H + h) * Q + q0 + r) * N + j] = Ss[r * SS + j];
    }
  }

  // ---- phase 2: scores -> basis weights r (in place)
  if (MODE == 0) {
    // per-row shift m = max(0, max_j S_j): the reference uses none (gibbs:248); it cancels exactly
    // between numerator and normaliser and only keeps exp() finite.
    for (int r = warp; r < rows; r += ATTN_THREADS / 32) {
      float m = 0.f;
      for (int j = lane; j < N; j += 32) m = fmaxf(m, Ss[r * SS + j]);
      m = warp_max(m);
      if (lane == 0) mrow[r] = m;
    }
    __syncthreads();
    if (p.hist_part) {
      rect_hist_tile([&](int r, int j) { return Ss[r * SS + j]; }, mrow, rows, p.jb, p.tb, tiles, zrow, part);
      __syncthreads();
      float* dst = p.hist_part + ((size_t)v * (H * gridDim.x) + h * gridDim.x + qt) * (EDGES - 2);
      for (int i = tid; i < EDGES - 2; i += ATTN_THREADS) dst[i] = part[i];
    }
    issue_tile(ntile);                                   // first V tile loads while the weights are formed
    for (int r = warp; r < QT; r += ATTN_THREADS / 32) {
      if (r >= rows) {                                   // unused query rows contribute nothing in phase 3
        for (int j = lane; j < N; j += 32) Ss[r * SS + j] = 0.f;
        continue;
      }
      const float m = mrow[r];
      float z = 0.f;
      for (int j = lane; j < N; j += 32) {
        const float w = tabA[j] * expf(Ss[r * SS + j] - m);
        Ss[r * SS + j] = w;
        z += w;
      }
      z = warp_sum(z) + p.W_out * expf(-m);
      for (int j = lane; j < N; j += 32) Ss[r * SS + j] = Ss[r * SS + j] / z;
    }
  } else {
    issue_tile(ntile);
    for (int r = warp; r < QT; r += ATTN_THREADS / 32) {
      if (r >= rows) {
        for (int j = lane; j < N; j += 32) Ss[r * SS + j] = 0.f;
        continue;
      }
      // a = softmax(20 S)  (gauss:289)
      float m = -INFINITY;
      for (int j = lane; j < N; j += 32) m = fmaxf(m, 20.f * Ss[r * SS + j]);
      m = warp_max(m);
      float z = 0.f;
      for (int j = lane; j < N; j += 32) {
        const float e = expf(20.f * Ss[r * SS + j] - m);
        Ss[r * SS + j] = e;
        z += e;
      }
      z = warp_sum(z);
      // mu = a.mu_b ; E[t^2] = a.(mu_b^2 + sigma_b^2)  (gauss:290-291), fp64 accumulation of the
      // fp32 products, rounded to fp32 before the (cancelling) subtraction exactly where the
      // reference rounds its matmul results.
      double am = 0.0, a2 = 0.0;
      for (int j = lane; j < N; j += 32) {
        const float a = Ss[r * SS + j] / z;
        const float bm = tabA[j], bs = tabB[j];
        const float c2 = __fadd_rn(__fmul_rn(bm, bm), __fmul_rn(bs, bs));
        am += (double)a * (double)bm;
        a2 += (double)a * (double)c2;
      }
      am = warp_sum(am);
      a2 = warp_sum(a2);
      const float mu = (float)am;
      const float var = __fsub_rn((float)a2, __fmul_rn(mu, mu));
      if (lane == 0) {
        const size_t o = (size_t)v * H * Q + (size_t)h * Q + q0 + r;
        if (p.mu_out) p.mu_out[o] = mu;
        if (p.sd_out) p.sd_out[o] = sqrtf(var);
      }
      // canonical-parameter round trip of the reference (gauss:308-310 + ContinuousSoftmax forward)
      const float th0 = __fdiv_rn(mu, var);
      const float th1 = __fdiv_rn(-1.f, __fmul_rn(2.f, var));
      const float var_rt = __fdiv_rn(-0.5f, th1);
      const float mu_rt = __fmul_rn(th0, var_rt);
      for (int j = lane; j < N; j += 32) {
        const float bs = tabB[j];
        const float s = sqrtf(__fadd_rn(__fmul_rn(bs, bs), var_rt));
        const float zz = __fdiv_rn(mu_rt - tabA[j], s);
        const float phi = 0.3989422804014327f * expf(-0.5f * zz * zz);
        Ss[r * SS + j] = __fdiv_rn(phi, s);
      }
    }
  }
  __syncthreads();

  // ---- phase 3: ctx[q, :] = sum_j r[q, j] V_h[j, :]; lane = query row with all 64 output columns in
  //      registers, each warp owns 8 rows of every V tile (split-j), partial sums are reduced at the end
  float acc[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) acc[c] = 0.f;
  for (int t = ntile; t < 2 * ntile; ++t) {
    if (t + 1 < 2 * ntile) { issue_tile(t + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const float* tile = tiles + (t & 1) * (JT * DH);
    const int jb0 = (t - ntile) * JT + warp * 8;
#pragma unroll 2
    for (int jj = 0; jj < 8; ++jj) {
      const int j = jb0 + jj;
      const float w = (j < N) ? Ss[lane * SS + j] : 0.f;
      const float4* vr = reinterpret_cast<const float4*>(tile + (warp * 8 + jj) * DH);
#pragma unroll
      for (int c4 = 0; c4 < DH / 4; ++c4) {
        const float4 x = vr[c4];
        acc[4 * c4 + 0] = fmaf(w, x.x, acc[4 * c4 + 0]);
        acc[4 * c4 + 1] = fmaf(w, x.y, acc[4 * c4 + 1]);
        acc[4 * c4 + 2] = fmaf(w, x.z, acc[4 * c4 + 2]);
        acc[4 * c4 + 3] = fmaf(w, x.w, acc[4 * c4 + 3]);
      }
    }
    __syncthreads();
  }
  // cross-warp reduction through shared memory: red[warp][c4][q] as float4 (conflict-free both ways)
  float4* red = reinterpret_cast<float4*>(smem);
#pragma unroll
  for (int c4 = 0; c4 < DH / 4; ++c4)
    red[(warp * (DH / 4) + c4) * QT + lane] = make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]);
  __syncthreads();
  for (int f = tid; f < (DH / 4) * QT; f += ATTN_THREADS) {
    const int c4 = f / QT, qr = f - c4 * QT;
    float4 s = red[c4 * QT + qr];
#pragma unroll
    for (int w = 1; w < ATTN_THREADS / 32; ++w) f4_add(s, red[(w * (DH / 4) + c4) * QT + qr]);
    if (qr < rows) *reinterpret_cast<float4*>(p.ctx + ((size_t)v * Q + q0 + qr) * D + h * DH + 4 * c4) = s;
  }
}

static size_t attn_smem_bytes(int N) {
  size_t f = (size_t)attn_front_floats(N) + QT * (DH + 1) + 2 * (size_t)N + 2 * QT + 128;
  return f * sizeof(float);
}

template <int MODE>
static int launch_attn(const AttnParams& p, int Bv, cudaStream_t stream, const char* name) {
  const size_t smem = attn_smem_bytes(p.N);
  LTM_REQUIRE(smem <= 227 * 1024, "%s: num_basis=%d needs %zu B of shared memory (> 227 KB)", name, p.N, smem);
  static PerDevice pd = {};
  if (int rc = kernel_setup(cont_attn_kernel<MODE>, smem, pd, nullptr)) return rc;
  dim3 grid((p.Q + QT - 1) / QT, p.H, Bv);
  cont_attn_kernel<MODE><<<grid, ATTN_THREADS, smem, stream>>>(p);
  LTM_CHECK_LAUNCH(name);
  return 0;
}

}  // namespace ltm

extern "C" int ltm_cont_attn_rect(const float* q, const float* KV, const float* W, float W_out,
                                  const int32_t* jb, const float* tb, float* ctx, float* scores_out,
                                  float* hist_part, int Bv, int Q, int N, int H, int d, void* stream) {
  using namespace ltm;
  LTM_REQUIRE(q && KV && W && ctx, "cont_attn_rect: null pointer");
  LTM_REQUIRE(hist_part == nullptr || (jb && tb), "cont_attn_rect: histogram requested without edge tables");
  LTM_REQUIRE(d == DH, "cont_attn_rect: head_size=%d unsupported (kernel is specialised for %d)", d, DH);
  LTM_REQUIRE(Bv > 0 && Bv <= 65535 && Q > 0 && N > 0 && H > 0 && H <= 65535, "cont_attn_rect: bad shape");
  LTM_REQUIRE(aligned16(q) && aligned16(KV) && aligned16(ctx), "cont_attn_rect: 16-byte alignment");
  AttnParams p{};
  p.q = q; p.KV = KV; p.tabA = W; p.tabB = nullptr; p.W_out = W_out; p.jb = jb; p.tb = tb; p.ctx = ctx;
  p.scores_out = scores_out; p.hist_part = hist_part; p.Q = Q; p.N = N; p.H = H;
  return launch_attn<0>(p, Bv, (cudaStream_t)stream, "cont_attn_rect");
}

extern "C" int ltm_cont_attn_gauss(const float* q, const float* KV, const float* basis_mu,
                                   const float* basis_sigma, float* ctx, float* scores_out, float* mu_out,
                                   float* sd_out, int Bv, int Q, int N, int H, int d, void* stream) {
  using namespace ltm;
  LTM_REQUIRE(q && KV && basis_mu && basis_sigma && ctx, "cont_attn_gauss: null pointer");
  LTM_REQUIRE(d == DH, "cont_attn_gauss: head_size=%d unsupported (kernel is specialised for %d)", d, DH);
  LTM_REQUIRE(Bv > 0 && Bv <= 65535 && Q > 0 && N > 0 && H > 0 && H <= 65535, "cont_attn_gauss: bad shape");
  LTM_REQUIRE(aligned16(q) && aligned16(KV) && aligned16(ctx), "cont_attn_gauss: 16-byte alignment");
  AttnParams p{};
  p.q = q; p.KV = KV; p.tabA = basis_mu; p.tabB = basis_sigma; p.ctx = ctx; p.scores_out = scores_out;
  p.mu_out = mu_out; p.sd_out = sd_out; p.Q = Q; p.N = N; p.H = H;
  return launch_attn<1>(p, Bv, (cudaStream_t)stream, "cont_attn_gauss");
}
