// Shared device routine: sticky histogram of a tile of score rows (used by sticky.cu and attn.cu).
#pragma once
#include "common.cuh"

namespace ltm {

constexpr int EDGES = LTM_STICKY_EDGES;   // 129 edges -> 128 intervals

// ------------------------------------------------------------------------------------------------
// Rect variant histogram for one tile of query rows held in shared memory.
//   E_i  = exp(z(tb_i) - m)            z = S[row, jb_i]  (0 where jb_i < 0: no basis is active)
//   Z    = trapz(E, tb)                (compute_probability :248)
//   p_i += dt_{i+1} (E_{i+1} + E_{i+2}) / (2 Z),  i = 0..126   (cumulative_trapezoid + diff :201-202,
//                                                              interval 0 is dropped upstream)
// `Srow(r, j)` reads the score of row r / basis j; `mrow[r]` is any per-row shift (it cancels).
// Eb: [rows][EDGES+1] scratch, Zb: [rows] scratch.  All threads of the CTA must call this.
// The sum over rows runs in a fixed order -> bit-reproducible.
// ------------------------------------------------------------------------------------------------
template <class ScoreAt>
__device__ __forceinline__ void rect_hist_tile(ScoreAt Srow, const float* mrow, int rows, const int32_t* __restrict__ jb,
                               const float* __restrict__ tb, float* Eb, float* Zb, float* out127) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int ES = EDGES + 1;
  for (int t = tid; t < rows * EDGES; t += nt) {
    const int r = t / EDGES, i = t - r * EDGES;
    const int j = jb[i];
    const float z = (j >= 0) ? Srow(r, j) : 0.f;
    Eb[r * ES + i] = expf(z - mrow[r]);
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  for (int r = warp; r < rows; r += nw) {
    float acc = 0.f;
    for (int i = lane; i < EDGES - 1; i += 32)
      acc += (tb[i + 1] - tb[i]) * (Eb[r * ES + i] + Eb[r * ES + i + 1]) * 0.5f;
    acc = warp_sum(acc);
    if (lane == 0) Zb[r] = acc;
  }
  __syncthreads();
  for (int i = tid; i < EDGES - 2; i += nt) {
    const float dt = tb[i + 2] - tb[i + 1];
    float acc = 0.f;
    for (int r = 0; r < rows; ++r) {
      const float z = Zb[r];
      acc += dt * (Eb[r * ES + i + 1] / z + Eb[r * ES + i + 2] / z) * 0.5f;
    }
    out127[i] = acc;
  }
}

}  // namespace ltm
