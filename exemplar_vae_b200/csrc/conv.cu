// K4 — convolution support for convhvae_2level / single_conv (sm_100a).
//
//   GatedConv2d / Conv2d  utils/nn.py:72-114,  weight-normed convs + ELU + Upsample  models/fully_conv.py:12-81
//
// Round-1 design: activations are NHWC; a convolution is  im2col (this file, HBM-bound gather)  ->  the
// dense-layer GEMM of K3 (tcgen05 3xTF32 with the fused gate / bias / activation epilogue)  ->  col2im
// (gather form, no atomics) for the input gradient.  The patch matrix is [N*OH*OW, kh*kw*C] with the
// channel index fastest, so both the gather and the GEMM operand loads are coalesced; weights are
// viewed as [Cout, kh*kw*Cin] by the host layer.  An im2col-free (implicit-GEMM, TMA im2col mode)
// kernel is the planned replacement once the models are parity-green.
#include <algorithm>

#include "common.cuh"

namespace exvae {
namespace {

inline int ew_blocks(long long n) { return (int)std::min<long long>((n + 255) / 256, 148LL * 32); }

// col[m, (ky*kw + kx)*C + c] = x[n, oy*s + ky - p, ox*s + kx - p, c]   (0 outside the image)
__global__ void __launch_bounds__(256) im2col_nhwc_kernel(const float* __restrict__ x, int N, int H, int W, int C,
                                                          int kh, int kw, int stride, int pad, int OH, int OW,
                                                          float* __restrict__ col) {
  const long long K = (long long)kh * kw * C;
  const long long total = (long long)N * OH * OW * K;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long m = e / K;
    const int k = (int)(e - m * K);
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
                                                          float* __restrict__ dx) {
  const long long K = (long long)kh * kw * C;
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
}  // namespace exvae

using namespace exvae;

extern "C" int exvae_im2col_nhwc(const float* x, int N, int H, int W, int C, int kh, int kw, int stride, int pad,
                                 float* col, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(x && col && N > 0 && H > 0 && W > 0 && C > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0);
  const int OH = (H + 2 * pad - kh) / stride + 1, OW = (W + 2 * pad - kw) / stride + 1;
  EXVAE_CHECK_ARG(OH > 0 && OW > 0);
  const long long total = (long long)N * OH * OW * kh * kw * C;
  im2col_nhwc_kernel<<<ew_blocks(total), 256, 0, as_stream(stream)>>>(x, N, H, W, C, kh, kw, stride, pad, OH, OW, col);
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_col2im_nhwc(const float* dcol, int N, int H, int W, int C, int kh, int kw, int stride, int pad,
                                 float* dx, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(dcol && dx && N > 0 && H > 0 && W > 0 && C > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0);
  const int OH = (H + 2 * pad - kh) / stride + 1, OW = (W + 2 * pad - kw) / stride + 1;
  EXVAE_CHECK_ARG(OH > 0 && OW > 0);
  const long long total = (long long)N * H * W * C;
  col2im_nhwc_kernel<<<ew_blocks(total), 256, 0, as_stream(stream)>>>(dcol, N, H, W, C, kh, kw, stride, pad, OH, OW, dx);
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
