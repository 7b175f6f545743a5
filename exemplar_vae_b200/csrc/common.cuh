// Shared device/host helpers for the exvae_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/exvae_b200.h"

namespace exvae {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kLog2Pi = 1.8378770664093453f;  // log(2*pi)

#define EXVAE_CHECK_ARG(cond) \
  do {                        \
    if (!(cond)) return EXVAE_ERR_INVALID_ARG; \
  } while (0)

// Launch-error check that never synchronises (safe during CUDA-graph capture).
#define EXVAE_RETURN_LAST_ERROR()              \
  do {                                         \
    cudaError_t e__ = cudaGetLastError();      \
    return e__ == cudaSuccess ? EXVAE_OK : (int)e__; \
  } while (0)

#define EXVAE_CUDA(call)                        \
  do {                                          \
    cudaError_t e__ = (call);                   \
    if (e__ != cudaSuccess) return (int)e__;    \
  } while (0)

__host__ __device__ constexpr inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ constexpr inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline cudaStream_t as_stream(exvae_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();  // cached cudaDevAttrMultiProcessorCount of the current device (api.cu)

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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

// Online log-sum-exp state in base 2: value = 2^m * s.  Merge is associative/commutative.
__device__ __forceinline__ void lse2_merge(float& m, float& s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  const float ms = (mn == -INFINITY) ? 0.f : mn;
  s = s * ex2_approx(m - ms) + s2 * ex2_approx(m2 - ms);
  m = mn;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + bulk async copy (TMA engine, SASS: UBLKCP / SYNCS) -----------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(__cvta_generic_to_global(src_gmem)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

}  // namespace exvae
