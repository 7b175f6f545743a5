// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/mma_rate tools/mma_rate.cu -lcuda   (run on a B200)
// Micro-benchmark: how fast can ONE thread issue tcgen05.mma.kind::tf32, and how long does the tensor pipe take per
// instruction, for N = 64 / 128 / 256 with A from shared memory or from tensor memory?
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../exemplar_vae_b200/csrc/tc_common.cuh"
using namespace exvae;
using namespace exvae::tc;
namespace exvae { int sm_count() { return 148; } }

template <int N, bool A_TM, int UNROLL>
__global__ void __launch_bounds__(128, 1) rate_kernel(unsigned long long* out, int iters) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc(128, N, false, false);
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
    long long t0 = clock64();
    unsigned long long g0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const uint64_t bd = umma_desc(sb + (u & 3) * 32, 16, 1024, 2);
        if (A_TM) umma_tf32_ts(tm, tm + 256 + 8 * (u & 3), bd, idesc, 1u);
        else umma_tf32(tm, umma_desc(sa + (u & 3) * 32, 16, 1024, 2), bd, idesc, 1u);
      }
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    unsigned long long g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    out[0] = t1 - t0; out[1] = t2 - t0; out[2] = g1 - g0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N, bool A_TM>
void run(const char* name) {
  unsigned long long* d; cudaMalloc(&d, 64);
  auto k = rate_kernel<N, A_TM, 12>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) { k<<<1, 128, 60000>>>(d, iters); cudaDeviceSynchronize(); }
  unsigned long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  const double n = iters * 12.0;
  printf("%-28s issue %.1f clk/MMA, issue+drain %.1f clk/MMA, %.1f ns/MMA -> %.0f flop/clk/SM (err %s)\n", name, h[0] / n, h[1] / n,
         h[2] / n, 2.0 * 128 * N * 8 / (h[1] / n), cudaGetErrorString(cudaGetLastError()));
  // all SMs at once
  k<<<148, 128, 60000>>>(d, iters); cudaDeviceSynchronize();
  cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  printf("%-28s (148 CTAs) issue+drain %.1f clk/MMA, %.1f ns/MMA\n", name, h[1] / n, h[2] / n);
  cudaFree(d);
}

int main() {
  run<64, false>("N=64  A smem");
  run<128, false>("N=128 A smem");
  run<256, false>("N=256 A smem");
  run<64, true>("N=64  A tmem");
  run<128, true>("N=128 A tmem");
  run<256, true>("N=256 A tmem");
  return 0;
}
