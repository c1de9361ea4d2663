// Shared helpers for the libinfltm kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/infltm.h"

namespace ltm {

void set_error(const char* fmt, ...);

#define LTM_REQUIRE(cond, ...)                  \
  do {                                          \
    if (!(cond)) {                              \
      ::ltm::set_error(__VA_ARGS__);            \
      return -1;                                \
    }                                           \
  } while (0)

#define LTM_CHECK_LAUNCH(name)                                                        \
  do {                                                                                \
    cudaError_t _e = cudaGetLastError();                                              \
    if (_e != cudaSuccess) {                                                          \
      ::ltm::set_error("%s: launch failed: %s", name, cudaGetErrorString(_e));        \
      return -2;                                                                      \
    }                                                                                 \
  } while (0)

#define LTM_CUDA(call)                                                                \
  do {                                                                                \
    cudaError_t _e = (call);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ::ltm::set_error("%s failed: %s", #call, cudaGetErrorString(_e));               \
      return -3;                                                                      \
    }                                                                                 \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// The dynamic-shared-memory opt-in of a kernel and the SM count are properties of a DEVICE, so their one-time setup
// is cached per device ordinal (a process may drive several GPUs: a Q-former sharded with device_map).
constexpr int LTM_MAX_DEVICES = 64;
struct PerDevice {
  size_t smem[LTM_MAX_DEVICES];      // largest dynamic shared memory size opted into so far
  int num_sms[LTM_MAX_DEVICES];
};
template <class Kern>
static inline int kernel_setup(Kern kern, size_t smem_bytes, PerDevice& pd, int* num_sms) {
  int dev = 0;
  LTM_CUDA(cudaGetDevice(&dev));
  LTM_REQUIRE(dev >= 0 && dev < LTM_MAX_DEVICES, "device ordinal %d out of range", dev);
  if (smem_bytes > pd.smem[dev]) {
    LTM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    pd.smem[dev] = smem_bytes;
  }
  if (pd.num_sms[dev] == 0)
    LTM_CUDA(cudaDeviceGetAttribute(&pd.num_sms[dev], cudaDevAttrMultiProcessorCount, dev));
  if (num_sms) *num_sms = pd.num_sms[dev];
  return 0;
}

#ifdef __CUDACC__
// 128-bit streaming load: read-only path, no L1 allocation, L2 evict-first (data is touched once).
__device__ __forceinline__ float4 ldg_stream(const float4* p, uint64_t policy) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p), "l"(policy));
  return r;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ldg_nc(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void f4_add(float4& a, const float4& b) {
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}
// the same four IEEE additions as two packed instructions (sm_100 add.rn.f32x2): half the issue slots of the
// accumulation in the streaming kernels
__device__ __forceinline__ void f4_add_packed(float4& a, const float4& b) {
  unsigned long long a0, a1, b0, b1;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a0) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(a1) : "f"(a.z), "f"(a.w));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b0) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b1) : "f"(b.z), "f"(b.w));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a0) : "l"(b0));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a1) : "l"(b1));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(a0));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a.z), "=f"(a.w) : "l"(a1));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif

}  // namespace ltm
