// G1/G2 -- Gaussian RBF design matrix and the ridge operator  G = F^T (F F^T + ridge I)^-1
// (long_term_attention.py:70-86, basis_functions.py:158-164).
//
// The operator depends only on (positions, basis, ridge), never on data, so it is solved once per
// shape and cached by the host.  The system is badly conditioned in fp32 (cond 5e5 at N=256, 2e7 at
// N=512: the reference's own fp32 `.inverse()` is off by 1-60 % against exact arithmetic), far beyond
// what TF32 tensor cores could resolve, so the one-off solve runs in fp64 on the CUDA cores
// (SYRK -> Cholesky -> triangular inverse -> G); the per-chunk contractions that *use* G run on
// tcgen05 (gemm_tcgen05.cu).
#include "common.cuh"

namespace ltm {

// out[p, j] = N(t_p; mu_j, sigma_j^2) with the reference's fp32 operation order
__global__ void rbf_eval_kernel(const float* __restrict__ tvals, const int32_t* __restrict__ tidx,
                                const float* __restrict__ mu, const float* __restrict__ sigma,
                                float* __restrict__ out, long long ld, int P, int N) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  if (j >= N) return;
  const float t = tidx ? tvals[tidx[p]] : tvals[p];
  const float s = sigma[j];
  const float z = __fdiv_rn(t - mu[j], s);
  const float phi = 0.3989422804014327f * expf(-0.5f * (z * z));
  out[(size_t)p * ld + j] = __fdiv_rn(phi, s);
}

// F[j][p] in fp64
__global__ void design_f64_kernel(const float* __restrict__ pos, const float* __restrict__ mu,
                                  const float* __restrict__ sigma, double* __restrict__ F, int P, int N) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (p >= P) return;
  const double s = (double)sigma[j];
  const double z = ((double)pos[p] - (double)mu[j]) / s;
  F[(size_t)j * P + p] = 0.3989422804014326779 * exp(-0.5 * z * z) / s;
}

// A = F F^T + ridge I   (32x32 output tile per CTA, 16x16 threads x 2x2 outputs)
__global__ void __launch_bounds__(256)
syrk_f64_kernel(const double* __restrict__ F, double* __restrict__ A, int P, int N, double ridge) {
  __shared__ double Fa[32][33];
  __shared__ double Fb[32][33];
  const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
  if (b0 > a0) return;                                   // lower triangle (+ mirrored below)
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[2][2] = {{0, 0}, {0, 0}};
  for (int p0 = 0; p0 < P; p0 += 32) {
    for (int f = threadIdx.x; f < 1024; f += 256) {
      const int r = f >> 5, c = f & 31;
      const int p = p0 + c;
      Fa[r][c] = (a0 + r < N && p < P) ? F[(size_t)(a0 + r) * P + p] : 0.0;
      Fb[r][c] = (b0 + r < N && p < P) ? F[(size_t)(b0 + r) * P + p] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < 32; ++c) {
      const double x0 = Fa[ty * 2][c], x1 = Fa[ty * 2 + 1][c];
      const double y0 = Fb[tx * 2][c], y1 = Fb[tx * 2 + 1][c];
      acc[0][0] += x0 * y0; acc[0][1] += x0 * y1; acc[1][0] += x1 * y0; acc[1][1] += x1 * y1;
    }
    __syncthreads();
  }
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) {
      const int a = a0 + ty * 2 + i, b = b0 + tx * 2 + j;
      if (a < N && b < N) {
        const double v = acc[i][j] + (a == b ? ridge : 0.0);
        A[(size_t)a * N + b] = v;
        A[(size_t)b * N + a] = v;
      }
    }
}

// In-place lower Cholesky, one CTA.  A (N x N, row-major) -> L in the lower triangle.
__global__ void __launch_bounds__(1024)
cholesky_f64_kernel(double* __restrict__ A, int N) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int k = 0; k < N; ++k) {
    __syncthreads();
    if (tid == 0) A[(size_t)k * N + k] = sqrt(A[(size_t)k * N + k]);
    __syncthreads();
    const double d = A[(size_t)k * N + k];
    for (int i = k + 1 + tid; i < N; i += nt) A[(size_t)i * N + k] /= d;
    __syncthreads();
    const int rem = N - k - 1;
    // trailing update of the lower triangle: (i, j), k < j <= i
    for (int f = tid; f < rem * rem; f += nt) {
      const int i = k + 1 + f / rem, j = k + 1 + f % rem;
      if (j <= i) A[(size_t)i * N + j] -= A[(size_t)i * N + k] * A[(size_t)j * N + k];
    }
  }
}

// X = L^-1 (lower triangular), one thread per column.
__global__ void tri_inverse_f64_kernel(const double* __restrict__ L, double* __restrict__ X, int N) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  for (int i = 0; i < c; ++i) X[(size_t)i * N + c] = 0.0;
  for (int i = c; i < N; ++i) {
    double s = (i == c) ? 1.0 : 0.0;
    for (int k = c; k < i; ++k) s -= L[(size_t)i * N + k] * X[(size_t)k * N + c];
    X[(size_t)i * N + c] = s / L[(size_t)i * N + i];
  }
}

// Ainv = X^T X
__global__ void xtx_f64_kernel(const double* __restrict__ X, double* __restrict__ Ainv, int N) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int a = blockIdx.y;
  if (b >= N) return;
  double s = 0.0;
  for (int i = max(a, b); i < N; ++i) s += X[(size_t)i * N + a] * X[(size_t)i * N + b];
  Ainv[(size_t)a * N + b] = s;
}

// G[p, j] = sum_i F[i, trim+p] Ainv[i, j]   (p < rows)
__global__ void __launch_bounds__(256)
apply_f64_kernel(const double* __restrict__ F, const double* __restrict__ Ainv, float* __restrict__ G,
                 float* __restrict__ GT, long long ldgt, int P, int N, int trim, int rows) {
  __shared__ double Fs[32][33];     // [i][p]
  __shared__ double As[32][33];     // [i][j]
  const int p0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  double acc[4] = {0, 0, 0, 0};
  for (int i0 = 0; i0 < N; i0 += 32) {
    for (int f = threadIdx.x; f < 1024; f += 256) {
      const int r = f >> 5, c = f & 31;
      Fs[r][c] = (i0 + r < N && p0 + c < rows) ? F[(size_t)(i0 + r) * P + trim + p0 + c] : 0.0;
      As[r][c] = (i0 + r < N && j0 + c < N) ? Ainv[(size_t)(i0 + r) * N + j0 + c] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      const double a = As[i][tx];
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[t] += Fs[i][ty * 4 + t] * a;
    }
    __syncthreads();
  }
  const int j = j0 + tx;
  for (int t = 0; t < 4; ++t) {
    const int p = p0 + ty * 4 + t;
    if (p < rows && j < N) {
      if (G) G[(size_t)p * N + j] = (float)acc[t];
      if (GT) GT[(size_t)j * ldgt + p] = (float)acc[t];
    }
  }
}

}  // namespace ltm

extern "C" int ltm_rbf_eval(const float* tvals, const int32_t* tidx, const float* basis_mu,
                            const float* basis_sigma, float* out, int64_t ld, int P, int N, void* stream) {
  using namespace ltm;
  LTM_REQUIRE(tvals && basis_mu && basis_sigma && out, "rbf_eval: null pointer");
  LTM_REQUIRE(P > 0 && P <= 65535 && N > 0 && ld >= N, "rbf_eval: bad shape P=%d N=%d ld=%lld", P, N, (long long)ld);
  dim3 grid((N + 127) / 128, P);
  rbf_eval_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(tvals, tidx, basis_mu, basis_sigma, out, ld, P, N);
  LTM_CHECK_LAUNCH("rbf_eval");
  return 0;
}

extern "C" int64_t ltm_ridge_workspace_doubles(int P, int N) {
  return (int64_t)N * P + 3ll * N * N;
}

extern "C" int ltm_ridge_solve(const float* positions, int P, int trim, int rows, const float* basis_mu,
                               const float* basis_sigma, int N, double ridge, float* G, float* GT,
                               int64_t ldgt, double* workspace, void* stream) {
  using namespace ltm;
  LTM_REQUIRE(positions && basis_mu && basis_sigma && workspace && (G || GT), "ridge_solve: null pointer");
  LTM_REQUIRE(N > 0 && N <= 2048 && P > 0 && trim >= 0 && rows > 0 && trim + rows <= P,
              "ridge_solve: bad shape N=%d P=%d trim=%d rows=%d", N, P, trim, rows);
  LTM_REQUIRE(GT == nullptr || ldgt >= rows, "ridge_solve: ldgt=%lld < rows=%d", (long long)ldgt, rows);
  cudaStream_t st = (cudaStream_t)stream;
  double* F = workspace;
  double* A = F + (size_t)N * P;
  double* X = A + (size_t)N * N;
  double* Ainv = X + (size_t)N * N;
  design_f64_kernel<<<dim3((P + 255) / 256, N), 256, 0, st>>>(positions, basis_mu, basis_sigma, F, P, N);
  LTM_CHECK_LAUNCH("ridge_solve/design");
  const int nt = (N + 31) / 32;
  syrk_f64_kernel<<<dim3(nt, nt), 256, 0, st>>>(F, A, P, N, ridge);
  LTM_CHECK_LAUNCH("ridge_solve/syrk");
  cholesky_f64_kernel<<<1, 1024, 0, st>>>(A, N);
  LTM_CHECK_LAUNCH("ridge_solve/cholesky");
  tri_inverse_f64_kernel<<<(N + 63) / 64, 64, 0, st>>>(A, X, N);
  LTM_CHECK_LAUNCH("ridge_solve/tri_inverse");
  xtx_f64_kernel<<<dim3((N + 127) / 128, N), 128, 0, st>>>(X, Ainv, N);
  LTM_CHECK_LAUNCH("ridge_solve/xtx");
  apply_f64_kernel<<<dim3((rows + 31) / 32, (N + 31) / 32), 256, 0, st>>>(F, Ainv, G, GT, (long long)ldgt, P, N, trim, rows);
  LTM_CHECK_LAUNCH("ridge_solve/apply");
  return 0;
}
