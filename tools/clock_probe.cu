// %globaltimer vs clock64 calibration (build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/clock_probe tools/clock_probe.cu)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void busy(unsigned long long* out, int mode) {
  unsigned long long g0 = gtimer(), c0 = clock64();
  if (mode == 0) { float x = threadIdx.x; for (int i = 0; i < 200000; ++i) x = x * 1.0001f + 0.5f; if (x == 123.f) out[9] = 1; }
  else { for (int i = 0; i < 50; ++i) __nanosleep(1000); }
  unsigned long long c1 = clock64(), g1 = gtimer();
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = c1 - c0; out[blockIdx.x * 2 + 1] = g1 - g0; }
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 4096); unsigned long long h[4];
  for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 3; ++rep) {
    busy<<<1, 32>>>(d, mode); cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("mode %d: cycles %llu ns %llu -> %.3f GHz\n", mode, h[0], h[1], (double)h[0] / h[1]);
  }
  busy<<<148 * 4, 256>>>(d, 0); cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("full grid busy: cycles %llu ns %llu -> %.3f GHz\n", h[0], h[1], (double)h[0] / h[1]);
  return 0;
}
