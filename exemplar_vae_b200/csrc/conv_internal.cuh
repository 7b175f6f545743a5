// Internal interface of the convolution support kernels (conv.cu), used by the conv entry points in gemm.cu.
#pragma once
#include <cuda_runtime.h>

namespace exvae {

// patch matrix col [N*OH*OW][ldk] (columns >= kh*kw*C are zero) and its adjoint
int conv_im2col(const float* x, int N, int H, int W, int C, int kh, int kw, int stride, int pad, int OH, int OW, int ldk,
                float* col, cudaStream_t st);
int conv_col2im(const float* dcol, int N, int H, int W, int C, int kh, int kw, int stride, int pad, int OH, int OW,
                int ldk, float* dx, cudaStream_t st);
// dW[co][ci][kh][kw] (+)= sum_s part[s][co_cat][t*cpad + ci] (rows < oseg -> dW0, rest -> dW1) and db from cs [S2][ncat]
int conv_unpack_wgrad(const float* part, int S, int ncat, int Kp, int cpad, int oseg, int Cin, int KH, int KW, float* dW0,
                      float* dW1, const float* cs, int S2, float* db0, float* db1, int accumulate, cudaStream_t st);

// dst [N][Hp][Wp][C] = src [N][h][w][C] placed at (oy0, ox0), zero elsewhere (C % 4 == 0, 16-byte aligned): the
// zero-padded frames of the implicit weight gradient
int conv_pad_frame(const float* src, int N, int h, int w, int C, int Hp, int Wp, int oy0, int ox0, float* dst,
                   cudaStream_t st);

}  // namespace exvae
