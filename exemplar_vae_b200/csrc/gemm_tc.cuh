// Internal interface of the tcgen05 3xTF32 GEMM (gemm_tc.cu), used by the K3 entry points in gemm.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace exvae {

enum TcEpi { TC_BIAS_ACT = 0, TC_GATED = 1, TC_PLAIN = 2, TC_SPLITK = 3, TC_LSE = 4, TC_PW = 5 };

// Exemplar-prior epilogues (K1 for D >= 64, prior_lse.cu): the GEMM computes S = Z'.M'^T = base-2 logits
// (rows = latents, columns = exemplars; operands augmented so that no bias term is needed).
//   TC_LSE: per row and column tile, the (max, sum 2^(S-max), #masked) partial of the log-sum-exp -> part[M][ntn][4]
//   TC_PW : W = g_row * 2^(S - lse2_row) (0 for leave-one-out pairs) stored as w [M][ldw] AND transposed wt [N][ldwt]
struct TcPriorEpi {
  const long long* cidx;   // [N] dataset index per column (null = no mask)
  const long long* zidx;   // [M] dataset index per row
  const float* g;          // [M]   TC_PW: upstream gradient
  const float* lse2;       // [M]   TC_PW: base-2 row log-sum
  float* part;             // TC_LSE
  float* w; int ldw;       // TC_PW
  float* wt; int ldwt;     // TC_PW
};

// Implicit-GEMM convolution: the A operand is never materialised.  Row m of A is output pixel m (NHWC, pixel index
// = (n*OH + oh)*OW + ow) and its K axis is (tap, channel) with channels padded to a multiple of 32 per tap:
//   A[m, (kh*KW + kw)*cpad + c] = x[n, oh*stride + kh - pad, ow*stride + kw - pad, c]      (0 outside the image)
// A 128-row tile is bn images x bh output rows x all OW pixels of each row, so that every k-block (one tap, 32
// channels) is ONE 4-D TMA box of the activation tensor; B holds the packed weights [N][taps*cpad] (K-major).
struct TcConv {
  const float* x;          // [N][H][W][C] fp32, 16-byte aligned, (C*4) % 16 == 0
  int N, H, W, C;
  int KH, KW, stride, pad;
  int OH, OW;
  int bh, bn;              // tile = bn images x bh rows x OW pixels <= 128 rows; bh | OH; bn > 1 only if bh == OH
};
// tile geometry for an output of OH x OW pixels; false if the shape cannot be tiled (OW > 128)
bool tc_conv_tiling(int OH, int OW, int* bh, int* bn);

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
  TcPriorEpi prior;    // TC_LSE / TC_PW
  const TcConv* conv;  // non-null: A is the implicit patch matrix of this convolution (a, a_rows, a_cols unused; a_mn = false)
  // Implicit weight gradient of a stride-1 convolution (b_mn, dwc_cin > 0): B is the zero-padded NHWC input viewed as
  // [pixels][dwc_cin]; output column n = tap*dwc_cin + ci reads B rows SHIFTED by the tap, k + (tap / dwc_kw) * dwc_wp +
  // tap % dwc_kw (dwc_wp = padded image width), so no patch matrix exists.  dwc_cin % 32 == 0; b_cols == dwc_cin.
  int dwc_cin, dwc_kw, dwc_wp;
};

bool tc_enabled();                       // sm_100 device, driver entry point found, not disabled by EXVAE_GEMM=simt
bool tc_dims_ok(int rows_pitch_elems);   // TMA needs 16-byte row pitches
void tc_set_trace(unsigned long long* buf);   // debug: 8 u64 per CTA of the NEXT launches (null = off)
unsigned long long* tc_take_trace(size_t words);   // debug: current trace segment (or null), then advance by `words`
int tc_gemm_launch(const TcGemm& g, cudaStream_t st);
int tc_gemm_ntn(const TcGemm& g);        // number of column tiles tc_gemm_launch will use for this problem (TC_LSE partials)
// out[0..n) = w0, out[n..2n) = w1 (the two weight tensors of a gated layer as one [2*O, K] operand); n % 4 == 0
int tc_concat2(const float* w0, const float* w1, size_t n, float* out, cudaStream_t st);

}  // namespace exvae
