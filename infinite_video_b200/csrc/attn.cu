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
// One CTA per (32-query tile, head, video); K_h / V_h tiles of 64 basis rows stream through shared
// memory (coalesced 128-bit loads of 256-byte row segments); each lane owns one query row for
// S = q K^T (K rows are warp-broadcast from shared memory) and each thread 8 output columns for r V.
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

__host__ __device__ inline int attn_tile_floats() {
  const int a = JT * DH, b = QT * (EDGES + 1);
  return a > b ? a : b;
}

template <int MODE>  // 0 = rect, 1 = gauss
__global__ void __launch_bounds__(ATTN_THREADS)
cont_attn_kernel(const AttnParams p) {
  extern __shared__ __align__(16) float smem[];
  const int N = p.N, Q = p.Q, H = p.H, D = H * DH;
  const int SS = N + 1;                                  // padded row stride of the score tile
  float* Ss = smem;                                      // [QT][N+1]
  float* tile = Ss + ((QT * SS + 3) & ~3);               // [JT][DH]  (aliased by the histogram scratch)
  float* qs = tile + attn_tile_floats();                 // [QT][DH+1]
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
    if (MODE == 1) {
      const float m = p.tabA[j], s = p.tabB[j];
      tabB[j] = s;
      (void)m;
    }
  }
  __syncthreads();
  float qreg[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) qreg[c] = qs[lane * (DH + 1) + c];

  // ---- phase 1: S[q, j] = q_h . K_h[j]
  for (int j0 = 0; j0 < N; j0 += JT) {
    const int jn = min(JT, N - j0);
    __syncthreads();
    for (int f = tid; f < JT * (DH / 4); f += ATTN_THREADS) {
      const int r = f / (DH / 4), c4 = f - r * (DH / 4);
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < jn) val = ldg_nc(reinterpret_cast<const float4*>(KVv + (size_t)(j0 + r) * 2 * D + h * DH) + c4);
      reinterpret_cast<float4*>(tile)[f] = val;
    }
    __syncthreads();
    for (int jj = warp; jj < jn; jj += ATTN_THREADS / 32) {
      const float4* kr = reinterpret_cast<const float4*>(tile + jj * DH);
      float acc = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < DH / 4; ++c4) {
        const float4 kk = kr[c4];
        acc = fmaf(qreg[4 * c4 + 0], kk.x, acc);
        acc = fmaf(qreg[4 * c4 + 1], kk.y, acc);
        acc = fmaf(qreg[4 * c4 + 2], kk.z, acc);
        acc = fmaf(qreg[4 * c4 + 3], kk.w, acc);
      }
      Ss[lane * SS + j0 + jj] = acc;
    }
  }
  __syncthreads();

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
      rect_hist_tile([&](int r, int j) { return Ss[r * SS + j]; }, mrow, rows, p.jb, p.tb, tile, zrow, part);
      __syncthreads();
      float* dst = p.hist_part + ((size_t)v * (H * gridDim.x) + h * gridDim.x + qt) * (EDGES - 2);
      for (int i = tid; i < EDGES - 2; i += ATTN_THREADS) dst[i] = part[i];
    }
    for (int r = warp; r < rows; r += ATTN_THREADS / 32) {
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
    for (int r = warp; r < rows; r += ATTN_THREADS / 32) {
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

  // ---- phase 3: ctx[q, :] = sum_j r[q, j] V_h[j, :]
  const int qr = tid >> 3;            // 0..31
  const int dg = tid & 7;             // 8 columns each
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int j0 = 0; j0 < N; j0 += JT) {
    const int jn = min(JT, N - j0);
    __syncthreads();
    for (int f = tid; f < JT * (DH / 4); f += ATTN_THREADS) {
      const int r = f / (DH / 4), c4 = f - r * (DH / 4);
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < jn)
        val = ldg_nc(reinterpret_cast<const float4*>(KVv + (size_t)(j0 + r) * 2 * D + D + h * DH) + c4);
      reinterpret_cast<float4*>(tile)[f] = val;
    }
    __syncthreads();
    const float* rr = Ss + qr * SS + j0;
    for (int jj = 0; jj < jn; ++jj) {
      const float w = rr[jj];
      const float4 v0 = reinterpret_cast<const float4*>(tile + jj * DH + dg * 8)[0];
      const float4 v1 = reinterpret_cast<const float4*>(tile + jj * DH + dg * 8)[1];
      acc[0] = fmaf(w, v0.x, acc[0]); acc[1] = fmaf(w, v0.y, acc[1]);
      acc[2] = fmaf(w, v0.z, acc[2]); acc[3] = fmaf(w, v0.w, acc[3]);
      acc[4] = fmaf(w, v1.x, acc[4]); acc[5] = fmaf(w, v1.y, acc[5]);
      acc[6] = fmaf(w, v1.z, acc[6]); acc[7] = fmaf(w, v1.w, acc[7]);
    }
  }
  if (qr < rows) {
    float4* dst = reinterpret_cast<float4*>(p.ctx + ((size_t)v * Q + q0 + qr) * D + h * DH + dg * 8);
    dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

static size_t attn_smem_bytes(int N) {
  size_t f = ((size_t)(QT * (N + 1) + 3) & ~(size_t)3) + attn_tile_floats() + QT * (DH + 1) + 2 * (size_t)N +
             2 * QT + 128;
  return f * sizeof(float);
}

template <int MODE>
static int launch_attn(const AttnParams& p, int Bv, cudaStream_t stream, const char* name) {
  const size_t smem = attn_smem_bytes(p.N);
  LTM_REQUIRE(smem <= 227 * 1024, "%s: num_basis=%d needs %zu B of shared memory (> 227 KB)", name, p.N, smem);
  static size_t configured[2] = {0, 0};
  if (smem > configured[MODE]) {
    LTM_CUDA(cudaFuncSetAttribute(cont_attn_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[MODE] = smem;
  }
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
