// K4 — convolution support for convhvae_2level / single_conv (sm_100a).
//
//   GatedConv2d / Conv2d  utils/nn.py:72-114,  weight-normed convs + ELU + Upsample  models/fully_conv.py:12-81
//
// Activations are NHWC.  A convolution with >= 16 input channels is an IMPLICIT GEMM on the tensor cores (gemm_tc.cu,
// TcConv): the patch matrix never exists, every k-block (one filter tap x 32 channels) is one 4-D TMA box of the
// activation tensor, zero padding comes from TMA's out-of-bounds fill; the gate / bias / activation epilogue is the
// dense layers'.  The same kernel gives the input gradient of stride-1 convolutions (a convolution of the
// pre-activation gradient with the flipped, transposed filters).  This file holds the support kernels: filter
// packing ([Cout][tap][channel padded to 32]) and its adjoint (weight-gradient unpack + split-K reduction + bias
// column sums), and the materialising im2col / col2im pair that remains for (a) layers with < 16 input channels
// (K = 9 / 27 / 49: the patch matrix is no bigger than the layer's output), (b) the weight gradient (patches are
// recomputed in the backward, never saved) and (c) the input gradient of stride-2 layers.  Patch rows are padded to a
// multiple of 4 floats so that every GEMM runs on tcgen05 (no FMA-pipe fallback).
#include <algorithm>

#include "common.cuh"

namespace exvae {
namespace {

inline int ew_blocks(long long n) { return (int)std::min<long long>((n + 255) / 256, 148LL * 32); }

// col[m, (ky*kw + kx)*C + c] = x[n, oy*s + ky - p, ox*s + kx - p, c]   (0 outside the image)
__global__ void __launch_bounds__(256) im2col_nhwc_kernel(const float* __restrict__ x, int N, int H, int W, int C,
                                                          int kh, int kw, int stride, int pad, int OH, int OW,
                                                          int ldk, float* __restrict__ col) {
  const int Kreal = kh * kw * C;
  const long long K = ldk;                       // row pitch >= kh*kw*C (padding columns are zero)
  const long long total = (long long)N * OH * OW * K;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long m = e / K;
    const int k = (int)(e - m * K);
    if (k >= Kreal) {
      col[e] = 0.f;
      continue;
    }
    const int c = k % C, kk = k / C;
    const int kx = kk % kw, ky = kk / kw;
    const int ox = (int)(m % OW);
    const long long t = m / OW;
    const int oy = (int)(t % OH), n = (int)(t / OH);
    const int iy = oy * stride + ky - pad, ix = ox * stride + kx - pad;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[(((long long)n * H + iy) * W + ix) * C + c];
    col[e] = v;
  }
}

// dx[n, iy, ix, c] = sum over (ky,kx) with (iy + p - ky) % s == 0 ... of dcol[m(n,oy,ox), (ky*kw+kx)*C + c]
__global__ void __launch_bounds__(256) col2im_nhwc_kernel(const float* __restrict__ dcol, int N, int H, int W, int C,
                                                          int kh, int kw, int stride, int pad, int OH, int OW,
                                                          int ldk, float* __restrict__ dx) {
  const long long K = ldk;
  const long long total = (long long)N * H * W * C;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    long long t = e / C;
    const int ix = (int)(t % W);
    t /= W;
    const int iy = (int)(t % H), n = (int)(t / H);
    float a = 0.f;
    for (int ky = 0; ky < kh; ++ky) {
      const int ty = iy + pad - ky;
      if (ty < 0 || ty % stride) continue;
      const int oy = ty / stride;
      if (oy >= OH) continue;
      for (int kx = 0; kx < kw; ++kx) {
        const int tx = ix + pad - kx;
        if (tx < 0 || tx % stride) continue;
        const int ox = tx / stride;
        if (ox >= OW) continue;
        const long long m = ((long long)n * OH + oy) * OW + ox;
        a += dcol[m * K + (long long)(ky * kw + kx) * C + c];
      }
    }
    dx[e] = a;
  }
}

// Filter packing.  mode 0 (forward operand, rows = output channels):
//     out[(row_off + co) * Kp + t * cpad + ci] = w[co][ci][kh][kw],  t = kh*KW + kw
// mode 1 (operand of the stride-1 input gradient, rows = INPUT channels, filters flipped):
//     out[ci * Kp + t' * cpad + (row_off + co)] = w[co][ci][KH-1-kh'][KW-1-kw'],  t' = kh'*KW + kw'
// (`out` is zeroed first: channel / K padding columns must be 0)
__global__ void __launch_bounds__(256) pack_conv_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int KH,
                                                               int KW, int cpad, int Kp, int mode, int row_off,
                                                               float* __restrict__ out) {
  const int total = Cout * Cin * KH * KW;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int kw = e % KW;
    int t = e / KW;
    const int kh = t % KH;
    t /= KH;
    const int ci = t % Cin, co = t / Cin;
    if (mode == 0)
      out[(size_t)(row_off + co) * Kp + (kh * KW + kw) * cpad + ci] = w[e];
    else
      out[(size_t)ci * Kp + ((KH - 1 - kh) * KW + (KW - 1 - kw)) * cpad + row_off + co] = w[e];
  }
}

// Adjoint of the mode-0 packing + split-K reduction: dW[co][ci][kh][kw] (+)= sum_s part[s][co_cat][t*cpad + ci];
// rows co_cat < oseg go to dW0, the rest to dW1 (gated layers: h and g filters).  The last blocks reduce the staging
// kernel's column sums cs [S2][ncat] into the bias gradients.
__global__ void __launch_bounds__(256) unpack_conv_wgrad_kernel(const float* __restrict__ part, int S, int ncat, int Kp,
                                                                int cpad, int oseg, int Cin, int KH, int KW,
                                                                float* __restrict__ dW0, float* __restrict__ dW1,
                                                                int nred, const float* __restrict__ cs, int S2,
                                                                float* __restrict__ db0, float* __restrict__ db1,
                                                                int accumulate) {
  if ((int)blockIdx.x < nred) {
    const int per = Cin * KH * KW;
    const int total = ncat * per;
    const size_t plane = (size_t)ncat * Kp;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += nred * blockDim.x) {
      const int co = e / per;
      int r = e - co * per;
      const int kw = r % KW;
      r /= KW;
      const int kh = r % KH, ci = r / KH;
      const size_t src = (size_t)co * Kp + (kh * KW + kw) * cpad + ci;
      float a = 0.f;
      for (int s = 0; s < S; ++s) a += part[(size_t)s * plane + src];
      float* dst = co < oseg ? dW0 + (size_t)co * per + (e - co * per) : dW1 + (size_t)(co - oseg) * per + (e - co * per);
      *dst = accumulate ? *dst + a : a;
    }
    return;
  }
  const int c = (blockIdx.x - nred) * 256 + threadIdx.x;
  if (c >= ncat) return;
  float* db = c < oseg ? db0 : db1;
  if (!db) return;
  float a = 0.f;
  for (int s = 0; s < S2; ++s) a += cs[(size_t)s * ncat + c];
  const int o = c < oseg ? c : c - oseg;
  db[o] = accumulate ? db[o] + a : a;
}

__global__ void __launch_bounds__(256) elu_fwd_kernel(const float* __restrict__ x, long long n, float* __restrict__ y) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float v = x[e];
    y[e] = v > 0.f ? v : expm1f(v);
  }
}
__global__ void __launch_bounds__(256) elu_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                      long long n, float* __restrict__ dx) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float v = y[e];                      // y = elu(x): x > 0 <=> y > 0 ; d/dx = 1 or y + 1
    dx[e] = dy[e] * (v > 0.f ? 1.f : v + 1.f);
  }
}

// nearest-neighbour 2x upsample, NHWC
__global__ void __launch_bounds__(256) up2_fwd_kernel(const float* __restrict__ x, int N, int H, int W, int C,
                                                      float* __restrict__ y) {
  const long long total = (long long)N * 2 * H * 2 * W * C;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    long long t = e / C;
    const int ox = (int)(t % (2 * W));
    t /= (2 * W);
    const int oy = (int)(t % (2 * H)), n = (int)(t / (2 * H));
    y[e] = x[(((long long)n * H + oy / 2) * W + ox / 2) * C + c];
  }
}
__global__ void __launch_bounds__(256) up2_bwd_kernel(const float* __restrict__ dy, int N, int H, int W, int C,
                                                      float* __restrict__ dx) {
  const long long total = (long long)N * H * W * C;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    long long t = e / C;
    const int ix = (int)(t % W);
    t /= W;
    const int iy = (int)(t % H), n = (int)(t / H);
    const long long b = (((long long)n * 2 * H + 2 * iy) * 2 * W + 2 * ix) * C + c;
    const long long rs = (long long)2 * W * C;
    dx[e] = (dy[b] + dy[b + C]) + (dy[b + rs] + dy[b + rs + C]);
  }
}

}  // namespace

int conv_im2col(const float* x, int N, int H, int W, int C, int kh, int kw, int stride, int pad, int OH, int OW, int ldk,
                float* col, cudaStream_t st) {
  const long long total = (long long)N * OH * OW * ldk;
  im2col_nhwc_kernel<<<ew_blocks(total), 256, 0, st>>>(x, N, H, W, C, kh, kw, stride, pad, OH, OW, ldk, col);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? EXVAE_OK : (int)e;
}
int conv_col2im(const float* dcol, int N, int H, int W, int C, int kh, int kw, int stride, int pad, int OH, int OW,
                int ldk, float* dx, cudaStream_t st) {
  const long long total = (long long)N * H * W * C;
  col2im_nhwc_kernel<<<ew_blocks(total), 256, 0, st>>>(dcol, N, H, W, C, kh, kw, stride, pad, OH, OW, ldk, dx);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? EXVAE_OK : (int)e;
}
int conv_unpack_wgrad(const float* part, int S, int ncat, int Kp, int cpad, int oseg, int Cin, int KH, int KW, float* dW0,
                      float* dW1, const float* cs, int S2, float* db0, float* db1, int accumulate, cudaStream_t st) {
  const int nred = std::max(1, std::min(ew_blocks((long long)ncat * Cin * KH * KW), 148 * 4));
  const int ncs = (db0 || db1) ? ceil_div(ncat, 256) : 0;
  unpack_conv_wgrad_kernel<<<nred + ncs, 256, 0, st>>>(part, S, ncat, Kp, cpad, oseg, Cin, KH, KW, dW0, dW1 ? dW1 : dW0,
                                                       nred, cs, S2, db0, db1, accumulate);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? EXVAE_OK : (int)e;
}

}  // namespace exvae

using namespace exvae;

extern "C" int exvae_conv_pack_weight(const float* w, int Cout, int Cin, int KH, int KW, int mode, int cpad, int Kp,
                                      int row_off, int zero_first, int rows_total, float* out, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(w && out && Cout > 0 && Cin > 0 && KH > 0 && KW > 0 && (mode == 0 || mode == 1) && cpad > 0 && Kp > 0);
  cudaStream_t st = as_stream(stream);
  if (zero_first) EXVAE_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)rows_total * Kp, st));
  const int total = Cout * Cin * KH * KW;
  pack_conv_weight_kernel<<<std::min(ceil_div(total, 256), 148 * 4), 256, 0, st>>>(w, Cout, Cin, KH, KW, cpad, Kp, mode,
                                                                                    row_off, out);
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_im2col_nhwc(const float* x, int N, int H, int W, int C, int kh, int kw, int stride, int pad,
                                 float* col, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(x && col && N > 0 && H > 0 && W > 0 && C > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0);
  const int OH = (H + 2 * pad - kh) / stride + 1, OW = (W + 2 * pad - kw) / stride + 1;
  EXVAE_CHECK_ARG(OH > 0 && OW > 0);
  const long long total = (long long)N * OH * OW * kh * kw * C;
  im2col_nhwc_kernel<<<ew_blocks(total), 256, 0, as_stream(stream)>>>(x, N, H, W, C, kh, kw, stride, pad, OH, OW,
                                                                      kh * kw * C, col);
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_col2im_nhwc(const float* dcol, int N, int H, int W, int C, int kh, int kw, int stride, int pad,
                                 float* dx, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(dcol && dx && N > 0 && H > 0 && W > 0 && C > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0);
  const int OH = (H + 2 * pad - kh) / stride + 1, OW = (W + 2 * pad - kw) / stride + 1;
  EXVAE_CHECK_ARG(OH > 0 && OW > 0);
  const long long total = (long long)N * H * W * C;
  col2im_nhwc_kernel<<<ew_blocks(total), 256, 0, as_stream(stream)>>>(dcol, N, H, W, C, kh, kw, stride, pad, OH, OW,
                                                                      kh * kw * C, dx);
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_elu_fwd(const float* x, int64_t n, float* y, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(x && y && n > 0);
  elu_fwd_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(x, n, y);
  EXVAE_RETURN_LAST_ERROR();
}
extern "C" int exvae_elu_bwd(const float* y, const float* dy, int64_t n, float* dx, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(y && dy && dx && n > 0);
  elu_bwd_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(y, dy, n, dx);
  EXVAE_RETURN_LAST_ERROR();
}
extern "C" int exvae_upsample2x_nhwc_fwd(const float* x, int N, int H, int W, int C, float* y, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(x && y && N > 0 && H > 0 && W > 0 && C > 0);
  up2_fwd_kernel<<<ew_blocks((long long)N * H * W * C * 4), 256, 0, as_stream(stream)>>>(x, N, H, W, C, y);
  EXVAE_RETURN_LAST_ERROR();
}
extern "C" int exvae_upsample2x_nhwc_bwd(const float* dy, int N, int H, int W, int C, float* dx,
                                         exvae_stream_t stream) {
  EXVAE_CHECK_ARG(dy && dx && N > 0 && H > 0 && W > 0 && C > 0);
  up2_bwd_kernel<<<ew_blocks((long long)N * H * W * C), 256, 0, as_stream(stream)>>>(dy, N, H, W, C, dx);
  EXVAE_RETURN_LAST_ERROR();
}
