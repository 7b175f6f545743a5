// NVLink 5 / NVSwitch exchanges of the range-sharded training step, written directly against peer memory (sm_100a).
//
// The reference is single-device; the B200 build shards the exemplar bank over the 8 GPUs of a box, which
// adds five small exchanges per step (latents + indices, LSE partials, row gradients, dz, parameter gradients).  They
// are all <= 5 MB, i.e. latency-bound: through NCCL (ring, LL protocol) each costs 18-48 us (measured on 8 x B200,
// tools/coll_probe.py).  Here every rank owns a SYMMETRIC buffer (same layout on every GPU, mapped into every peer and
// into one NVSwitch multicast address, set up by torch.distributed._symmetric_memory) and the kernels use
//   multimem.st          one store lands in all 8 GPUs                      (all-gather: each rank publishes its slot)
//   multimem.ld_reduce   one load returns the SUM over the 8 GPUs, reduced IN THE SWITCH   (all-reduce, reduce-scatter)
// plus a flag barrier over per-peer signal pads (release / acquire compare-and-swap at system scope).  No staging
// copies, no ring hops: an exchange is one kernel of a few CTAs.
#include "common.cuh"

namespace exvae {
namespace {

constexpr int MC_THREADS = 512;

__device__ __forceinline__ uint32_t cas_release_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}

// Barrier between the CTAs with the same `channel` on every rank.  pads[r] = signal pad of rank r (uint32 slots,
// [channel][world]).  Thread t < world raises slot [channel][rank] on peer t (waiting for it to be clear first) and
// then waits for, and clears, slot [channel][t] on its own pad: self-resetting, so back-to-back barriers are safe.
// The leading __syncthreads + release makes every earlier write of the CTA visible to the peers that pass the barrier.
__device__ __forceinline__ void mc_barrier(uint32_t* const* pads, int rank, int world, int channel) {
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int peer = threadIdx.x;
    uint32_t* put = pads[peer] + (size_t)channel * world + rank;
    while (cas_release_sys(put, 0u, 1u) != 0u) {
    }
    uint32_t* get = pads[rank] + (size_t)channel * world + peer;
    while (cas_acquire_sys(get, 1u, 0u) != 1u) {
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float4 mm_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void mm_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// In-place all-reduce (sum * scale) of n floats of the symmetric buffer: rank r reduces slice r in the switch and
// multicasts the result back.  n % (4 * world) == 0.
__global__ void __launch_bounds__(MC_THREADS) mc_allreduce_kernel(float* __restrict__ mc, uint32_t* const* pads,
                                                                  long long n, int rank, int world, int ch0,
                                                                  float scale) {
  mc_barrier(pads, rank, world, ch0 + blockIdx.x);              // every rank's contribution is in place
  const long long per = n / world, base = per * rank;
  for (long long i = 4ll * (blockIdx.x * (long long)blockDim.x + threadIdx.x); i < per;
       i += 4ll * gridDim.x * blockDim.x) {
    float4 v = mm_ld_reduce_add(mc + base + i);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    mm_st(mc + base + i, v);
  }
  mc_barrier(pads, rank, world, ch0 + blockIdx.x);              // every slice has landed everywhere
}

// All-gather: src (local, n floats, n % 4 == 0) -> slot `rank` of the symmetric [world][n] region on ALL ranks.
__global__ void __launch_bounds__(MC_THREADS) mc_allgather_kernel(const float* __restrict__ src, float* __restrict__ mc_dst,
                                                                  uint32_t* const* pads, long long n, int rank,
                                                                  int world, int ch0) {
  float* slot = mc_dst + n * rank;
  for (long long i = 4ll * (blockIdx.x * (long long)blockDim.x + threadIdx.x); i < n; i += 4ll * gridDim.x * blockDim.x)
    mm_st(slot + i, *reinterpret_cast<const float4*>(src + i));
  mc_barrier(pads, rank, world, ch0 + blockIdx.x);
}

// Reduce-scatter: every rank holds a partial [world][n] in the symmetric region; out (local, n floats) = sum over the
// ranks of slice `rank`.  The trailing barrier lets the region be overwritten afterwards.
__global__ void __launch_bounds__(MC_THREADS) mc_reduce_scatter_kernel(const float* __restrict__ mc_src,
                                                                       float* __restrict__ out, uint32_t* const* pads,
                                                                       long long n, int rank, int world, int ch0) {
  mc_barrier(pads, rank, world, ch0 + blockIdx.x);
  const float* slice = mc_src + n * rank;
  for (long long i = 4ll * (blockIdx.x * (long long)blockDim.x + threadIdx.x); i < n; i += 4ll * gridDim.x * blockDim.x)
    *reinterpret_cast<float4*>(out + i) = mm_ld_reduce_add(slice + i);
  mc_barrier(pads, rank, world, ch0 + blockIdx.x);
}

inline int mc_blocks(long long work_floats, int max_blocks) {
  const long long b = (work_floats / 4 + MC_THREADS - 1) / MC_THREADS;
  return (int)std::max<long long>(1, std::min<long long>(b, max_blocks));
}
inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace exvae

using namespace exvae;

extern "C" int exvae_mc_allreduce(float* mc_ptr, void* signal_pads_dev, int64_t n, int rank, int world, int channel0,
                                  int max_blocks, float scale, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(mc_ptr && signal_pads_dev && n > 0 && world > 1 && rank >= 0 && rank < world && max_blocks > 0);
  EXVAE_CHECK_ARG(n % (4 * world) == 0 && al16(mc_ptr));
  const int blocks = mc_blocks(n / world, max_blocks);
  mc_allreduce_kernel<<<blocks, MC_THREADS, 0, as_stream(stream)>>>(mc_ptr, static_cast<uint32_t* const*>(signal_pads_dev), n,
                                                                   rank, world, channel0, scale);
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_mc_allgather(const float* src, float* mc_dst, void* signal_pads_dev, int64_t n, int rank, int world,
                                  int channel0, int max_blocks, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(src && mc_dst && signal_pads_dev && n > 0 && world > 1 && rank >= 0 && rank < world && max_blocks > 0);
  EXVAE_CHECK_ARG(n % 4 == 0 && al16(src) && al16(mc_dst));
  mc_allgather_kernel<<<mc_blocks(n, max_blocks), MC_THREADS, 0, as_stream(stream)>>>(
      src, mc_dst, static_cast<uint32_t* const*>(signal_pads_dev), n, rank, world, channel0);
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_mc_reduce_scatter(const float* mc_src, float* out, void* signal_pads_dev, int64_t n, int rank,
                                       int world, int channel0, int max_blocks, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(mc_src && out && signal_pads_dev && n > 0 && world > 1 && rank >= 0 && rank < world && max_blocks > 0);
  EXVAE_CHECK_ARG(n % 4 == 0 && al16(mc_src) && al16(out));
  mc_reduce_scatter_kernel<<<mc_blocks(n, max_blocks), MC_THREADS, 0, as_stream(stream)>>>(
      mc_src, out, static_cast<uint32_t* const*>(signal_pads_dev), n, rank, world, channel0);
  EXVAE_RETURN_LAST_ERROR();
}
