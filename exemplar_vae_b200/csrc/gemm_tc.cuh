// Internal interface of the tcgen05 3xTF32 GEMM (gemm_tc.cu), used by the K3 entry points in gemm.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace exvae {

enum TcEpi { TC_BIAS_ACT = 0, TC_GATED = 1, TC_PLAIN = 2, TC_SPLITK = 3 };

// D[M,N] = A[M,K] * B[N,K]^T (+ epilogue).  Operands are plain fp32 row-major matrices [rows][cols] (16-byte aligned
// base and row pitch); the kernel splits them into tf32 hi/lo parts in shared memory.
//   a_mn == false: A is [M rows][K cols] (K contiguous)       "K-major"
//   a_mn == true : A is [K rows][M cols] (M contiguous)       "MN-major"   (same for B with N)
struct TcGemm {
  const float* a; int a_rows, a_cols; bool a_mn;
  const float* b; int b_rows, b_cols; bool b_mn;
  int M, N, K;
  int epi;
  int gated_O;        // TC_GATED: B rows [0,O) are the h weights, [O,2O) the g weights; N == O
  const float* bias0; const float* bias1;
  float* out0; float* out1; float* out2; int ldc;
  int act; float lo, hi;
  int splits, kchunk;  // TC_SPLITK: grid.z = splits, out0 = partial [splits][M][N]
};

bool tc_enabled();                       // sm_100 device, driver entry point found, not disabled by EXVAE_GEMM=simt
bool tc_dims_ok(int rows_pitch_elems);   // TMA needs 16-byte row pitches
void tc_set_trace(unsigned long long* buf);   // debug: 8 u64 per CTA of the NEXT launches (null = off)
unsigned long long* tc_take_trace(size_t words);   // debug: current trace segment (or null), then advance by `words`
int tc_gemm_launch(const TcGemm& g, cudaStream_t st);
// out[0..n) = w0, out[n..2n) = w1 (the two weight tensors of a gated layer as one [2*O, K] operand); n % 4 == 0
int tc_concat2(const float* w0, const float* w1, size_t n, float* out, cudaStream_t st);

}  // namespace exvae
