// Internal interface of the tcgen05 3xTF32 GEMM (gemm_tc.cu), used by the K3 entry points in gemm.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace exvae {

enum TcEpi { TC_BIAS_ACT = 0, TC_GATED = 1, TC_PLAIN = 2, TC_SPLITK = 3 };

// D[M,N] = A[M,K] * B[N,K]^T (+ epilogue).  Operands are hi/lo split planes: split[0] = tf32(x),
// split[1] = tf32(x - hi), each a row-major matrix [rows][cols].
//   a_mn == false: planes are [M rows][K cols] (K contiguous)       "K-major"
//   a_mn == true : planes are [K rows][M cols] (M contiguous)       "MN-major"   (same for B with N)
struct TcGemm {
  const float* a_split; int a_rows, a_cols; bool a_mn;
  const float* b_split; int b_rows, b_cols; bool b_mn;
  int M, N, K;
  int epi;
  int gated_O;        // TC_GATED: B plane rows [0,O) are the h weights, [O,2O) the g weights; N == O
  const float* bias0; const float* bias1;
  float* out0; float* out1; float* out2; int ldc;
  int act; float lo, hi;
  int splits, kchunk;  // TC_SPLITK: grid.z = splits, out0 = partial [splits][M][N]
};

bool tc_enabled();                       // sm_100 device, driver entry point found, not disabled by EXVAE_GEMM=simt
bool tc_dims_ok(int rows_pitch_elems);   // TMA needs 16-byte row pitches
int tc_gemm_launch(const TcGemm& g, cudaStream_t st);
// out[0..n) = tf32_rna(x), out[plane_stride .. plane_stride+n) = tf32_rna(x - hi)
int tc_split(const float* x, size_t n, float* out, size_t plane_stride, cudaStream_t st);

}  // namespace exvae
