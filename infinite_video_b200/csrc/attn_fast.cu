// Fast path of the continuous attention (R10/R11 + R6, G4) for num_basis in {64, 128, 256}.
//
// Same mathematics as attn.cu (see the header there); different data flow:
//   * the K/V projection GEMM stores the keys transposed per head, Kt[v][h][d][j] (gemm_tcgen05.cu epilogue),
//     so S = q K^T is an outer-product accumulation over d with both operands read as conflict-free 128-bit
//     shared-memory vectors: warp = 4 query rows, lane = NB/32 basis columns, 4 x NB/32 accumulators per thread;
//   * the row statistics (max, normaliser, mu / E[t^2]) are warp-shuffle reductions on registers;
//   * the weights r are written once, transposed and XOR-swizzled (Rt[j][q]), and ctx = r V is a second
//     outer-product accumulation over j: 8 warps = 4 j-quarters x 2 d-halves, lane = 8 queries x 4 columns;
//   * K and V tiles stream through a two-slot cp.async ring.
// Per CTA: 2 x 524k FMA, ~3 LDS.128 per 32 FMA, ~70 registers.
#include "common.cuh"
#include "rect_hist.cuh"

namespace ltm {

namespace fast {

constexpr int DH = 64;
constexpr int QT = 32;
constexpr int THREADS = 256;
constexpr int RING = 4096;             // floats per ring slot (16 KB)

struct Params {
  const float* q;        // [Bv,Q,D]
  const float* Kt;       // [Bv,H,64,NB]
  const float* V;        // [Bv,NB,ldv] (head h at column h*64)
  long long ldv;
  const float* tabA;     // rect: W[NB]        gauss: basis_mu[NB]
  const float* tabB;     //                    gauss: basis_sigma[NB]
  float W_out;
  const int32_t* jb;
  const float* tb;
  float* ctx;            // [Bv,Q,D]
  float* scores_out;     // optional [Bv,H,Q,NB]
  float* hist_part;      // rect: optional [Bv, H*q_tiles, 127]
  float* mu_out;
  float* sd_out;
  int Q, H;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

template <int NB>
__host__ __device__ constexpr int smem_floats(bool hist) {
  // ring[2][RING] | Qt[64][32] | Rt[NB][32] | mrow[32] zrow[32] part[128]
  // The histogram scratch Eb[32][130] aliases Rt (written only after the histogram) when it fits (NB >= 130),
  // otherwise it gets its own region.
  return 2 * RING + DH * QT + NB * QT + 32 + 32 + 128 + ((hist && NB * QT < QT * (EDGES + 1)) ? QT * (EDGES + 1) : 0);
}

template <int MODE, int NB>
__global__ void __launch_bounds__(THREADS, (MODE == 0) ? 3 : 2)
cont_attn_fast_kernel(const Params p) {
  static_assert(NB == 64 || NB == 128 || NB == 256, "fast path covers num_basis 64/128/256");
  constexpr int JPT = NB / 32;                 // basis columns per lane in phase 1
  constexpr int DCH = RING / NB;               // Kt rows (d) per ring slot
  constexpr int NST1 = DH / DCH;               // phase-1 stages
  constexpr int NST3 = NB / 64;                // phase-3 stages (64 basis rows x 64 columns)
  extern __shared__ __align__(16) float smem[];
  float* ring = smem;
  float* Qt = ring + 2 * RING;
  float* Rt = Qt + DH * QT;
  float* mrow = Rt + NB * QT;
  float* zrow = mrow + 32;
  float* part = zrow + 32;
  float* Eb = (NB * QT >= QT * (EDGES + 1)) ? Rt : part + 128;

  const int Q = p.Q, H = p.H, D = H * DH;
  const int qt = blockIdx.x, h = blockIdx.y, v = blockIdx.z;
  const int q0 = qt * QT;
  const int rows = min(QT, Q - q0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* Ktv = p.Kt + ((size_t)v * H + h) * DH * NB;
  const float* Vv = p.V + (size_t)v * NB * p.ldv + h * DH;

  auto issue_k = [&](int st) {                 // Kt rows [st*DCH, st*DCH+DCH): contiguous RING floats
    float* dst = ring + (st & 1) * RING;
    const float* src = Ktv + (size_t)st * RING;
#pragma unroll
    for (int i = 0; i < RING / 4 / THREADS; ++i) {
      const int f = tid + i * THREADS;
      cp_async16(dst + 4 * f, src + 4 * f);
    }
    cp_async_commit();
  };
  auto issue_v = [&](int st, int slot) {       // V rows [st*64, st*64+64) x 64 columns
    float* dst = ring + slot * RING;
#pragma unroll
    for (int i = 0; i < RING / 4 / THREADS; ++i) {
      const int f = tid + i * THREADS;
      const int r = f >> 4, c4 = f & 15;
      cp_async16(dst + r * DH + 4 * c4, Vv + (size_t)(st * 64 + r) * p.ldv + 4 * c4);
    }
    cp_async_commit();
  };
  issue_k(0);

  // ---- query tile, scaled by 1/sqrt(d) (gibbs:226), transposed: Qt[d][q]
  {
    const float inv_sqrt_d = 1.0f / sqrtf((float)DH);
    const float* qrow = p.q + ((size_t)v * Q + q0 + lane) * D + h * DH;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int d = warp * 8 + i;
      Qt[d * QT + lane] = (lane < rows) ? qrow[d] * inv_sqrt_d : 0.f;
    }
  }

  // ---- phase 1: S[4 warp + i][JPT lane + jj] accumulated over d
  float acc[4][JPT];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int jj = 0; jj < JPT; ++jj) acc[i][jj] = 0.f;
#pragma unroll 1
  for (int st = 0; st < NST1; ++st) {
    if (st + 1 < NST1) { issue_k(st + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const float* kst = ring + (st & 1) * RING + JPT * lane;
    const float* qst = Qt + (st * DCH) * QT + 4 * warp;
#pragma unroll 8
    for (int dd = 0; dd < DCH; ++dd) {
      const float4 qv = *reinterpret_cast<const float4*>(qst + dd * QT);
      float kv[JPT];
      if constexpr (JPT == 8) {
        const float4 a = *reinterpret_cast<const float4*>(kst + dd * NB);
        const float4 b = *reinterpret_cast<const float4*>(kst + dd * NB + 4);
        kv[0] = a.x; kv[1] = a.y; kv[2] = a.z; kv[3] = a.w; kv[4] = b.x; kv[5] = b.y; kv[6] = b.z; kv[7] = b.w;
      } else if constexpr (JPT == 4) {
        const float4 a = *reinterpret_cast<const float4*>(kst + dd * NB);
        kv[0] = a.x; kv[1] = a.y; kv[2] = a.z; kv[3] = a.w;
      } else {
        const float2 a = *reinterpret_cast<const float2*>(kst + dd * NB);
        kv[0] = a.x; kv[1] = a.y;
      }
#pragma unroll
      for (int jj = 0; jj < JPT; ++jj) {
        acc[0][jj] = fmaf(qv.x, kv[jj], acc[0][jj]);
        acc[1][jj] = fmaf(qv.y, kv[jj], acc[1][jj]);
        acc[2][jj] = fmaf(qv.z, kv[jj], acc[2][jj]);
        acc[3][jj] = fmaf(qv.w, kv[jj], acc[3][jj]);
      }
    }
    __syncthreads();
  }
  // the ring is idle now: start the first V stage in slot 1.  Slot 0 serves as the [32][NB] score scratch of
  // the histogram; for NB = 256 that scratch needs both slots, so the V stage is issued after the histogram.
  const bool want_hist = (MODE == 0) && (p.hist_part != nullptr);
  const bool early_v = !(want_hist && NB == 256);
  if (early_v) issue_v(0, 1);

  if (p.scores_out) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = 4 * warp + i;
      if (r < rows) {
        float* dst = p.scores_out + (((size_t)v * H + h) * Q + q0 + r) * NB + JPT * lane;
#pragma unroll
        for (int jj = 0; jj < JPT; ++jj) dst[jj] = acc[i][jj];
      }
    }
  }

  // ---- phase 2 (registers + warp shuffles): scores -> weights r, written transposed + swizzled to Rt
  float tA[JPT], tB[JPT];
#pragma unroll
  for (int jj = 0; jj < JPT; ++jj) {
    tA[jj] = __ldg(p.tabA + JPT * lane + jj);
    tB[jj] = (MODE == 1) ? __ldg(p.tabB + JPT * lane + jj) : 0.f;
  }
  if (MODE == 0) {
    const bool hist = want_hist;
    float* Ss = ring;                           // [QT][NB] score scratch for the histogram
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // per-row shift m = max(0, max_j S_j): the reference uses none (gibbs:248); it cancels exactly
      float m = 0.f;
#pragma unroll
      for (int jj = 0; jj < JPT; ++jj) m = fmaxf(m, acc[i][jj]);
      m = warp_max(m);
      if (hist) {
#pragma unroll
        for (int jj = 0; jj < JPT; ++jj) Ss[(4 * warp + i) * NB + JPT * lane + jj] = acc[i][jj];
        if (lane == 0) mrow[4 * warp + i] = m;
      }
      float z = 0.f;
#pragma unroll
      for (int jj = 0; jj < JPT; ++jj) {
        acc[i][jj] = tA[jj] * expf(acc[i][jj] - m);
        z += acc[i][jj];
      }
      z = warp_sum(z) + p.W_out * expf(-m);
#pragma unroll
      for (int jj = 0; jj < JPT; ++jj) acc[i][jj] = acc[i][jj] / z;
    }
    if (hist) {
      __syncthreads();
      rect_hist_tile([&](int r, int j) { return Ss[r * NB + j]; }, mrow, rows, p.jb, p.tb, Eb, zrow, part);
      __syncthreads();
      float* dst = p.hist_part + ((size_t)v * (H * gridDim.x) + h * gridDim.x + qt) * (EDGES - 2);
      for (int i = tid; i < EDGES - 2; i += THREADS) dst[i] = part[i];
      if (!early_v) issue_v(0, 1);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // a = softmax(20 S)  (gauss:289)
      float m = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < JPT; ++jj) m = fmaxf(m, 20.f * acc[i][jj]);
      m = warp_max(m);
      float z = 0.f;
#pragma unroll
      for (int jj = 0; jj < JPT; ++jj) {
        acc[i][jj] = expf(20.f * acc[i][jj] - m);
        z += acc[i][jj];
      }
      z = warp_sum(z);
      // mu = a.mu_b ; E[t^2] = a.(mu_b^2 + sigma_b^2)  (gauss:290-291): fp64 accumulation of the fp32 products,
      // rounded to fp32 before the cancelling subtraction exactly where the reference rounds its matmul results
      double am = 0.0, a2 = 0.0;
#pragma unroll
      for (int jj = 0; jj < JPT; ++jj) {
        const float a = acc[i][jj] / z;
        const float c2 = __fadd_rn(__fmul_rn(tA[jj], tA[jj]), __fmul_rn(tB[jj], tB[jj]));
        am += (double)a * (double)tA[jj];
        a2 += (double)a * (double)c2;
      }
      am = warp_sum(am);
      a2 = warp_sum(a2);
      const float mu = (float)am;
      const float var = __fsub_rn((float)a2, __fmul_rn(mu, mu));
      const int r = 4 * warp + i;
      if (lane == 0 && r < rows) {
        const size_t o = (size_t)v * H * Q + (size_t)h * Q + q0 + r;
        if (p.mu_out) p.mu_out[o] = mu;
        if (p.sd_out) p.sd_out[o] = sqrtf(var);
      }
      // canonical-parameter round trip of the reference (gauss:308-310 + ContinuousSoftmax forward)
      const float th0 = __fdiv_rn(mu, var);
      const float th1 = __fdiv_rn(-1.f, __fmul_rn(2.f, var));
      const float var_rt = __fdiv_rn(-0.5f, th1);
      const float mu_rt = __fmul_rn(th0, var_rt);
#pragma unroll
      for (int jj = 0; jj < JPT; ++jj) {
        const float s = sqrtf(__fadd_rn(__fmul_rn(tB[jj], tB[jj]), var_rt));
        const float zz = __fdiv_rn(mu_rt - tA[jj], s);
        acc[i][jj] = __fdiv_rn(0.3989422804014327f * expf(-0.5f * zz * zz), s);
      }
    }
  }
  // Rt[j][q]: row j holds the 32 query weights as 8 float4 chunks, chunk c stored at c ^ ((j >> 3) & 7)
#pragma unroll
  for (int jj = 0; jj < JPT; ++jj) {
    const int j = JPT * lane + jj;
    *reinterpret_cast<float4*>(Rt + j * QT + 4 * (warp ^ ((j >> 3) & 7))) =
        make_float4(acc[0][jj], acc[1][jj], acc[2][jj], acc[3][jj]);
  }

  // ---- phase 3: ctx[q, :] = sum_j r[q, j] V_h[j, :]   (warp: j-quarter jw, column half dh; lane: 8 q x 4 d)
  const int jw = warp & 3, dh = warp >> 2, qg = lane >> 3, dg = lane & 7;
  float o[8][4];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) o[a][b] = 0.f;
#pragma unroll 1
  for (int st = 0; st < NST3; ++st) {
    const int slot = (st + 1) & 1;
    if (st + 1 < NST3) { issue_v(st + 1, slot ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();                            // also orders the Rt writes before the first reads
    const float* vst = ring + slot * RING + dh * 32 + 4 * dg + (jw * 16) * DH;
    // the warp's 16 basis rows of this stage are two runs of 8 with a constant swizzle term each
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int jbase = st * 64 + jw * 16 + half * 8;
      const int sw = (jbase >> 3) & 7;
      const float* rp0 = Rt + jbase * QT + 4 * ((2 * qg) ^ sw);
      const float* rp1 = Rt + jbase * QT + 4 * ((2 * qg + 1) ^ sw);
      const float* vp = vst + (half * 8) * DH;
#pragma unroll
      for (int jl = 0; jl < 8; ++jl) {
        const float4 r0 = *reinterpret_cast<const float4*>(rp0 + jl * QT);
        const float4 r1 = *reinterpret_cast<const float4*>(rp1 + jl * QT);
        const float4 vv = *reinterpret_cast<const float4*>(vp + jl * DH);
        const float rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          o[a][0] = fmaf(rr[a], vv.x, o[a][0]);
          o[a][1] = fmaf(rr[a], vv.y, o[a][1]);
          o[a][2] = fmaf(rr[a], vv.z, o[a][2]);
          o[a][3] = fmaf(rr[a], vv.w, o[a][3]);
        }
      }
    }
    __syncthreads();
  }
  // reduce the four j-quarters through shared memory (the ring is dead): red[jw][q][64]
  float* red = ring;
#pragma unroll
  for (int a = 0; a < 8; ++a)
    *reinterpret_cast<float4*>(red + ((jw * QT + 8 * qg + a) * DH) + dh * 32 + 4 * dg) =
        make_float4(o[a][0], o[a][1], o[a][2], o[a][3]);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int f = tid + i * THREADS;            // float4 index over [32 q][16]
    const int qr = f >> 4, c4 = f & 15;
    float4 s = *reinterpret_cast<const float4*>(red + qr * DH + 4 * c4);
#pragma unroll
    for (int w = 1; w < 4; ++w) f4_add(s, *reinterpret_cast<const float4*>(red + (w * QT + qr) * DH + 4 * c4));
    if (qr < rows) *reinterpret_cast<float4*>(p.ctx + ((size_t)v * Q + q0 + qr) * D + h * DH + 4 * c4) = s;
  }
}

template <int MODE, int NB>
static int launch(const Params& p, int Bv, bool hist, cudaStream_t stream, const char* name) {
  const size_t smem = sizeof(float) * smem_floats<NB>(MODE == 0);
  static PerDevice pd = {};
  if (int rc = kernel_setup(cont_attn_fast_kernel<MODE, NB>, smem, pd, nullptr)) return rc;
  (void)hist;
  dim3 grid((p.Q + QT - 1) / QT, p.H, Bv);
  cont_attn_fast_kernel<MODE, NB><<<grid, THREADS, smem, stream>>>(p);
  LTM_CHECK_LAUNCH(name);
  return 0;
}

template <int MODE>
static int dispatch(const Params& p, int N, int Bv, bool hist, cudaStream_t stream, const char* name) {
  switch (N) {
    case 64: return launch<MODE, 64>(p, Bv, hist, stream, name);
    case 128: return launch<MODE, 128>(p, Bv, hist, stream, name);
    case 256: return launch<MODE, 256>(p, Bv, hist, stream, name);
    default: set_error("%s: the transposed-key fast path covers num_basis 64/128/256, got %d", name, N); return -1;
  }
}

}  // namespace fast
}  // namespace ltm

extern "C" int ltm_attn_fast_supported(int N, int d) { return (d == 64 && (N == 64 || N == 128 || N == 256)) ? 1 : 0; }

extern "C" int ltm_cont_attn_rect_t(const float* q, const float* Kt, const float* V, int64_t ldv, const float* W,
                                    float W_out, const int32_t* jb, const float* tb, float* ctx, float* scores_out,
                                    float* hist_part, int Bv, int Q, int N, int H, int d, void* stream) {
  using namespace ltm;
  LTM_REQUIRE(q && Kt && V && W && ctx, "cont_attn_rect_t: null pointer");
  LTM_REQUIRE(hist_part == nullptr || (jb && tb), "cont_attn_rect_t: histogram requested without edge tables");
  LTM_REQUIRE(ltm_attn_fast_supported(N, d), "cont_attn_rect_t: unsupported num_basis=%d / head_size=%d", N, d);
  LTM_REQUIRE(Bv > 0 && Bv <= 65535 && Q > 0 && H > 0 && H <= 65535 && ldv % 4 == 0, "cont_attn_rect_t: bad shape");
  LTM_REQUIRE(aligned16(q) && aligned16(Kt) && aligned16(V) && aligned16(ctx), "cont_attn_rect_t: 16-byte alignment");
  fast::Params p{};
  p.q = q; p.Kt = Kt; p.V = V; p.ldv = ldv; p.tabA = W; p.W_out = W_out; p.jb = jb; p.tb = tb; p.ctx = ctx;
  p.scores_out = scores_out; p.hist_part = hist_part; p.Q = Q; p.H = H;
  return fast::dispatch<0>(p, N, Bv, hist_part != nullptr, (cudaStream_t)stream, "cont_attn_rect_t");
}

extern "C" int ltm_cont_attn_gauss_t(const float* q, const float* Kt, const float* V, int64_t ldv,
                                     const float* basis_mu, const float* basis_sigma, float* ctx, float* scores_out,
                                     float* mu_out, float* sd_out, int Bv, int Q, int N, int H, int d, void* stream) {
  using namespace ltm;
  LTM_REQUIRE(q && Kt && V && basis_mu && basis_sigma && ctx, "cont_attn_gauss_t: null pointer");
  LTM_REQUIRE(ltm_attn_fast_supported(N, d), "cont_attn_gauss_t: unsupported num_basis=%d / head_size=%d", N, d);
  LTM_REQUIRE(Bv > 0 && Bv <= 65535 && Q > 0 && H > 0 && H <= 65535 && ldv % 4 == 0, "cont_attn_gauss_t: bad shape");
  LTM_REQUIRE(aligned16(q) && aligned16(Kt) && aligned16(V) && aligned16(ctx), "cont_attn_gauss_t: 16-byte alignment");
  fast::Params p{};
  p.q = q; p.Kt = Kt; p.V = V; p.ldv = ldv; p.tabA = basis_mu; p.tabB = basis_sigma; p.ctx = ctx;
  p.scores_out = scores_out; p.mu_out = mu_out; p.sd_out = sd_out; p.Q = Q; p.H = H;
  return fast::dispatch<1>(p, N, Bv, false, (cudaStream_t)stream, "cont_attn_gauss_t");
}
