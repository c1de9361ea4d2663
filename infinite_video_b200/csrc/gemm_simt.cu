// fp32 CUDA-core batched GEMM with the same argument block as the tcgen05 kernel.
// Role: on-device check kernel for the tensor-core path (tests, debugging; impl = 1).  It is not the
// product path: ltm_gemm defaults to the tcgen05 kernel in gemm_tcgen05.cu.
#include "common.cuh"

namespace ltm {

constexpr int ST = 64;   // tile
constexpr int SK = 16;

__global__ void __launch_bounds__(256)
gemm_simt_kernel(const ltm_gemm_args g) {
  __shared__ float As[SK][ST + 4];
  __shared__ float Bs[SK][ST + 4];
  const int b = blockIdx.z;
  const int m0 = blockIdx.y * ST, n0 = blockIdx.x * ST;
  const float* A = g.A + (size_t)b * g.strideA;
  const float* B1 = g.B + (size_t)b * g.strideB;
  const float* B2 = g.B2 ? g.B2 + (size_t)b * g.strideB2 : nullptr;
  float* C = g.C + (size_t)b * g.strideC;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < g.K; k0 += SK) {
    for (int f = threadIdx.x; f < SK * ST; f += 256) {
      int kk, mm;
      if (g.a_kmajor) { mm = f / SK; kk = f - mm * SK; } else { kk = f / ST; mm = f - kk * ST; }
      const int m = m0 + mm, k = k0 + kk;
      float val = 0.f;
      if (m < g.M && k < g.K) {
        if (g.a_group > 0) val = A[(size_t)(m / g.a_group) * g.a_group_stride + (size_t)(m % g.a_group) * g.lda + k];
        else val = g.a_kmajor ? A[(size_t)m * g.lda + k] : A[(size_t)k * g.lda + m];
      }
      As[kk][mm] = val;
    }
    for (int f = threadIdx.x; f < SK * ST; f += 256) {
      int kk, nn;
      if (g.b_kmajor) { nn = f / SK; kk = f - nn * SK; } else { kk = f / ST; nn = f - kk * ST; }
      const int n = n0 + nn, k = k0 + kk;
      float val = 0.f;
      if (n < g.Nc && k < g.K) {
        const bool seg2 = (B2 != nullptr) && (k >= g.K1);
        const float* Bp = seg2 ? B2 : B1;
        const size_t ld = seg2 ? g.ldb2 : g.ldb;
        const int kr = seg2 ? k - g.K1 : k;
        val = g.b_kmajor ? Bp[(size_t)n * ld + kr] : Bp[(size_t)kr * ld + n];
      }
      Bs[kk][nn] = val;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SK; ++kk) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < g.Nc) {
        const float val = acc[i][j] + (g.bias ? g.bias[(size_t)b * g.bias_stride + n] : 0.f);
        if (g.CT != nullptr && n < g.ct_cols)
          (g.CT + (size_t)b * g.strideC)[((size_t)(m / g.ct_group) * g.ct_cols + n) * g.ct_group + (m % g.ct_group)] = val;
        else
          C[(g.c_group > 0 ? (size_t)(m / g.c_group) * g.c_group_stride + (size_t)(m % g.c_group) * g.ldc
                           : (size_t)m * g.ldc) + n - (g.CT != nullptr ? g.ct_cols : 0)] = val;
      }
    }
  }
}

int gemm_simt_launch(const ltm_gemm_args& g, cudaStream_t stream) {
  LTM_REQUIRE(g.batch > 0 && g.batch <= 65535, "gemm(simt): batch=%d out of range", g.batch);
  dim3 grid((g.Nc + ST - 1) / ST, (g.M + ST - 1) / ST, g.batch);
  LTM_REQUIRE(grid.y <= 65535, "gemm(simt): M too large");
  gemm_simt_kernel<<<grid, 256, 0, stream>>>(g);
  LTM_CHECK_LAUNCH("gemm(simt)");
  return 0;
}

}  // namespace ltm
