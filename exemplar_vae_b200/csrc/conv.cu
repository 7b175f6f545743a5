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

// col[m, (ky*kw + kx)*C + c] = x[n, oy*s + ky - p, ox*s + kx - p, c]   (0 outside the image); row pitch ldk >= kh*kw*C,
// padding columns zero.  One thread per VW consecutive channels of one tap of one row (VW = 4 when C % 4 == 0), all
// index arithmetic in 32 bits: the kernel is a pure gather and must run at HBM speed.
template <int VW>
__global__ void __launch_bounds__(256) im2col_nhwc_kernel(const float* __restrict__ x, int N, int H, int W, int C,
                                                          int kh, int kw, int stride, int pad, int OH, int OW,
                                                          int ldk, float* __restrict__ col) {
  const unsigned cv = (unsigned)(C / VW);                 // channel groups per tap
  const unsigned per_row = (unsigned)(kh * kw) * cv;      // work items per row (the K padding is written by item 0)
  const unsigned long long total = (unsigned long long)N * OH * OW * per_row;
  const int Kreal = kh * kw * C;
  for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < total;
       e += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned m = (unsigned)(e / per_row);
    const unsigned it = (unsigned)(e - (unsigned long long)m * per_row);
    const unsigned kk = it / cv, c = (it - kk * cv) * VW;
    const unsigned ky = kk / (unsigned)kw, kx = kk - ky * (unsigned)kw;
    const unsigned t = m / (unsigned)OW, ox = m - t * (unsigned)OW;
    const unsigned n = t / (unsigned)OH, oy = t - n * (unsigned)OH;
    const int iy = (int)oy * stride + (int)ky - pad, ix = (int)ox * stride + (int)kx - pad;
    float* dst = col + (size_t)m * ldk + kk * C + c;
    const bool in = iy >= 0 && iy < H && ix >= 0 && ix < W;
    const float* src = x + (((size_t)n * H + iy) * W + ix) * C + c;
    if (VW == 4) {
      *reinterpret_cast<float4*>(dst) = in ? *reinterpret_cast<const float4*>(src) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      *dst = in ? *src : 0.f;
    }
    if (it == 0)
      for (int k = Kreal; k < ldk; ++k) col[(size_t)m * ldk + k] = 0.f;
  }
}

// dx[n, iy, ix, c] = sum over the taps (ky,kx) with (iy + p - ky) % s == 0, ... of dcol[m(n,oy,ox), (ky*kw+kx)*C + c]
// (gather form of the adjoint: no atomics, deterministic); one thread per VW channels of one input pixel
template <int VW>
__global__ void __launch_bounds__(256) col2im_nhwc_kernel(const float* __restrict__ dcol, int N, int H, int W, int C,
                                                          int kh, int kw, int stride, int pad, int OH, int OW,
                                                          int ldk, float* __restrict__ dx) {
  const unsigned cv = (unsigned)(C / VW);
  const unsigned long long total = (unsigned long long)N * H * W * cv;
  for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < total;
       e += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned pix = (unsigned)(e / cv);
    const unsigned c = (unsigned)(e - (unsigned long long)pix * cv) * VW;
    const unsigned t = pix / (unsigned)W, ix = pix - t * (unsigned)W;
    const unsigned n = t / (unsigned)H, iy = t - n * (unsigned)H;
    float a[VW];
#pragma unroll
    for (int k = 0; k < VW; ++k) a[k] = 0.f;
    for (int ky = 0; ky < kh; ++ky) {
      const int ty = (int)iy + pad - ky;
      if (ty < 0 || ty % stride) continue;
      const int oy = ty / stride;
      if (oy >= OH) continue;
      for (int kx = 0; kx < kw; ++kx) {
        const int tx = (int)ix + pad - kx;
        if (tx < 0 || tx % stride) continue;
        const int ox = tx / stride;
        if (ox >= OW) continue;
        const size_t m = ((size_t)n * OH + oy) * OW + ox;
        const float* src = dcol + m * ldk + (ky * kw + kx) * C + c;
        if (VW == 4) {
          const float4 v = *reinterpret_cast<const float4*>(src);
          a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
        } else {
          a[0] += *src;
        }
      }
    }
    float* dst = dx + (size_t)pix * C + c;
    if (VW == 4) *reinterpret_cast<float4*>(dst) = make_float4(a[0], a[1], a[2], a[3]);
    else *dst = a[0];
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
    // walk the PACKED layout (coalesced reads of the split-K partials); the scattered writes hit the small filter tensor
    const int taps = KH * KW, per = Cin * taps;
    const int total = ncat * taps * Cin;
    const size_t plane = (size_t)ncat * Kp;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += nred * blockDim.x) {
      const int ci = e % Cin;
      const int r = e / Cin;
      const int t = r % taps, co = r / taps;
      const size_t src = (size_t)co * Kp + t * cpad + ci;
      float a = 0.f;
      for (int s = 0; s < S; ++s) a += part[(size_t)s * plane + src];
      const int o = ci * taps + t;                          // [ci][kh][kw] inside one filter
      float* dst = co < oseg ? dW0 + (size_t)co * per + o : dW1 + (size_t)(co - oseg) * per + o;
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


namespace {
// one thread per 16-byte group of the padded frame
__global__ void __launch_bounds__(256) pad_frame_kernel(const float4* __restrict__ src, int N, int h, int w, int C4, int Hp,
                                                        int Wp, int oy0, int ox0, float4* __restrict__ dst) {
  const unsigned long long total = (unsigned long long)N * Hp * Wp * C4;
  for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < total;
       e += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned c = (unsigned)(e % C4);
    const unsigned long long pix = e / C4;
    const unsigned x = (unsigned)(pix % Wp);
    const unsigned long long t = pix / Wp;
    const unsigned y = (unsigned)(t % Hp), n = (unsigned)(t / Hp);
    const int sy = (int)y - oy0, sx = (int)x - ox0;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sy >= 0 && sy < h && sx >= 0 && sx < w) v = src[(((size_t)n * h + sy) * w + sx) * C4 + c];
    dst[e] = v;
  }
}
}  // namespace

int conv_pad_frame(const float* src, int N, int h, int w, int C, int Hp, int Wp, int oy0, int ox0, float* dst,
                   cudaStream_t st) {
  if ((C & 3) || (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15)) return EXVAE_ERR_INVALID_ARG;
  const long long total = (long long)N * Hp * Wp * (C / 4);
  pad_frame_kernel<<<ew_blocks(total), 256, 0, st>>>(reinterpret_cast<const float4*>(src), N, h, w, C / 4, Hp, Wp, oy0, ox0,
                                                     reinterpret_cast<float4*>(dst));
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? EXVAE_OK : (int)e;
}

int conv_im2col(const float* x, int N, int H, int W, int C, int kh, int kw, int stride, int pad, int OH, int OW, int ldk,
                float* col, cudaStream_t st) {
  const bool v4 = C % 4 == 0 && ldk % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(col) & 15) == 0;
  const long long total = (long long)N * OH * OW * kh * kw * (C / (v4 ? 4 : 1));
  if (v4) im2col_nhwc_kernel<4><<<ew_blocks(total), 256, 0, st>>>(x, N, H, W, C, kh, kw, stride, pad, OH, OW, ldk, col);
  else im2col_nhwc_kernel<1><<<ew_blocks(total), 256, 0, st>>>(x, N, H, W, C, kh, kw, stride, pad, OH, OW, ldk, col);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? EXVAE_OK : (int)e;
}
int conv_col2im(const float* dcol, int N, int H, int W, int C, int kh, int kw, int stride, int pad, int OH, int OW,
                int ldk, float* dx, cudaStream_t st) {
  const bool v4 = C % 4 == 0 && ldk % 4 == 0 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(dcol) & 15) == 0;
  const long long total = (long long)N * H * W * (C / (v4 ? 4 : 1));
  if (v4) col2im_nhwc_kernel<4><<<ew_blocks(total), 256, 0, st>>>(dcol, N, H, W, C, kh, kw, stride, pad, OH, OW, ldk, dx);
  else col2im_nhwc_kernel<1><<<ew_blocks(total), 256, 0, st>>>(dcol, N, H, W, C, kh, kw, stride, pad, OH, OW, ldk, dx);
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
  return conv_im2col(x, N, H, W, C, kh, kw, stride, pad, OH, OW, kh * kw * C, col, as_stream(stream));
}

extern "C" int exvae_col2im_nhwc(const float* dcol, int N, int H, int W, int C, int kh, int kw, int stride, int pad,
                                 float* dx, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(dcol && dx && N > 0 && H > 0 && W > 0 && C > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0);
  const int OH = (H + 2 * pad - kh) / stride + 1, OW = (W + 2 * pad - kw) / stride + 1;
  EXVAE_CHECK_ARG(OH > 0 && OW > 0);
  return conv_col2im(dcol, N, H, W, C, kh, kw, stride, pad, OH, OW, kh * kw * C, dx, as_stream(stream));
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
