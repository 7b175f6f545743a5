// K3 — dense layers of the encoder/decoder (sm_100a), fp32 FMA-pipe GEMM with fused epilogues.
//
//   GatedDense  utils/nn.py:44-69    out = (x Wh^T + bh) * sigmoid(x Wg^T + bg)
//   NonLinear / nn.Linear  utils/nn.py:29-41   out = act(x W^T + b), act in {none, sigmoid, hardtanh}
// and their backward passes (dx, dW, db).
//
// One kernel template covers the four operand layouts the forward/backward need:
//   A "KC": A[m][k] with k contiguous (activations, gradients)      A "XC": A[k][m] with m contiguous (dY^T)
//   B "KC": B[n][k] with k contiguous (nn.Linear weights, forward)  B "XC": B[k][n] with n contiguous (W for dx, x for dW)
// Operand B (and XC-A) may be split in two segments so the h- and g-branch weights of a gated
// layer are read in place (the reference keeps them as two nn.Linear modules, and the
// state_dict key names must survive).  In the gated forward a 128-wide tile holds 64 h-columns
// and the 64 matching g-columns, so h*sigmoid(g) is formed in registers.
//
// CTA tile 128x128x16, 256 threads, 8x8 outputs per thread (two 4x4 quadrant pairs), smem
// double buffering with register prefetch: one __syncthreads per k-step.
// Precision: fp32 operands and accumulation (the 1e-4 ELBO parity bar rules out single-pass
// tf32/bf16 tensor-core operands; a 3xTF32 tcgen05 path is the planned upgrade).
#include <algorithm>

#include "common.cuh"
#include "conv_internal.cuh"
#include "gemm_tc.cuh"

namespace exvae {
namespace {

constexpr int GM = 128, GN = 128, GK = 16, GP = GM + 4;
enum { L_KC = 0, L_XC = 1 };
enum { EPI_BIAS_ACT = 0, EPI_GATED = 1, EPI_PLAIN = 2, EPI_SPLITK = 3 };

struct GemmP {
  const float* A;
  int lda;
  const float* B0;
  const float* B1;
  int ldb;
  int bseg;  // XC-B: rows k >= bseg come from B1.  KC-B (non gated): unused
  int M, N, K;
  int kchunk;  // split-K: blockIdx.z covers [z*kchunk, min(K,(z+1)*kchunk))
  const float* bias0;
  const float* bias1;
  float* out0;
  float* out1;
  float* out2;
  int ldc;
  int act;
  float lo, hi;
  int O;       // gated forward: number of output columns
  int a_vec;   // A rows are 16-byte aligned and lda % 4 == 0
  int b_vec;
  int c_vec;   // output rows are 16-byte aligned and ldc % 4 == 0
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// load 4 consecutive elements starting at p[i0], valid while index < limit
__device__ __forceinline__ float4 load4(const float* __restrict__ p, int i0, int limit, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p == nullptr) return v;
  if (vec && i0 + 3 < limit) return *reinterpret_cast<const float4*>(p + i0);
  if (i0 < limit) v.x = p[i0];
  if (i0 + 1 < limit) v.y = p[i0 + 1];
  if (i0 + 2 < limit) v.z = p[i0 + 2];
  if (i0 + 3 < limit) v.w = p[i0 + 3];
  return v;
}

template <int AL, int BL, int EPI>
__global__ void __launch_bounds__(256, 2) sgemm_kernel(const GemmP p) {
  __shared__ __align__(16) float As[2][GK][GP];
  __shared__ __align__(16) float Bs[2][GK][GP];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * GM;
  const int n0 = (EPI == EPI_GATED) ? blockIdx.x * (GN / 2) : blockIdx.x * GN;
  const int kbeg = blockIdx.z * p.kchunk;
  const int kend = min(p.K, kbeg + p.kchunk);

  // per-thread load coordinates
  const int kc_r = tid >> 2, kc_k = (tid & 3) * 4;   // KC: rows kc_r, kc_r+64 ; k offset kc_k
  const int xc_k = tid >> 5, xc_x = (tid & 31) * 4;  // XC: k rows xc_k, xc_k+8 ; x offset xc_x

  const float* a_row[2] = {nullptr, nullptr};
  if (AL == L_KC) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = m0 + kc_r + 64 * h;
      a_row[h] = m < p.M ? p.A + (size_t)m * p.lda : nullptr;
    }
  }
  const float* b_row[2] = {nullptr, nullptr};
  if (BL == L_KC) {
    if (EPI == EPI_GATED) {
      const int j = n0 + kc_r;
      b_row[0] = j < p.O ? p.B0 + (size_t)j * p.ldb : nullptr;
      b_row[1] = j < p.O ? p.B1 + (size_t)j * p.ldb : nullptr;
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = n0 + kc_r + 64 * h;
        b_row[h] = n < p.N ? p.B0 + (size_t)n * p.ldb : nullptr;
      }
    }
  }

  float4 ra[2], rb[2];
  auto fetch = [&](int k0) {
    if (AL == L_KC) {
#pragma unroll
      for (int h = 0; h < 2; ++h) ra[h] = load4(a_row[h], k0 + kc_k, kend, p.a_vec);
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = k0 + xc_k + 8 * h;
        ra[h] = load4(k < kend ? p.A + (size_t)k * p.lda : nullptr, m0 + xc_x, p.M, p.a_vec);
      }
    }
    if (BL == L_KC) {
#pragma unroll
      for (int h = 0; h < 2; ++h) rb[h] = load4(b_row[h], k0 + kc_k, kend, p.b_vec);
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = k0 + xc_k + 8 * h;
        const float* row = nullptr;
        if (k < kend) row = k < p.bseg ? p.B0 + (size_t)k * p.ldb : p.B1 + (size_t)(k - p.bseg) * p.ldb;
        rb[h] = load4(row, n0 + xc_x, p.N, p.b_vec);
      }
    }
  };
  auto stash = [&](int buf) {
    if (AL == L_KC) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = kc_r + 64 * h;
        As[buf][kc_k + 0][r] = ra[h].x;
        As[buf][kc_k + 1][r] = ra[h].y;
        As[buf][kc_k + 2][r] = ra[h].z;
        As[buf][kc_k + 3][r] = ra[h].w;
      }
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) *reinterpret_cast<float4*>(&As[buf][xc_k + 8 * h][xc_x]) = ra[h];
    }
    if (BL == L_KC) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = kc_r + 64 * h;
        Bs[buf][kc_k + 0][r] = rb[h].x;
        Bs[buf][kc_k + 1][r] = rb[h].y;
        Bs[buf][kc_k + 2][r] = rb[h].z;
        Bs[buf][kc_k + 3][r] = rb[h].w;
      }
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) *reinterpret_cast<float4*>(&Bs[buf][xc_k + 8 * h][xc_x]) = rb[h];
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = (kend > kbeg) ? (kend - kbeg + GK - 1) / GK : 0;
  if (nk > 0) {
    fetch(kbeg);
    stash(0);
  }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) fetch(kbeg + (kt + 1) * GK);
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) stash(buf ^ 1);
    __syncthreads();
  }

  // ------------------------------------------------------------------ epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
    if (EPI == EPI_GATED) {
      const int j0 = n0 + tx * 4;
      float o[4], hh[4], ss[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = j0 + j;
        const float bh = (col < p.O && p.bias0) ? p.bias0[col] : 0.f;
        const float bg = (col < p.O && p.bias1) ? p.bias1[col] : 0.f;
        hh[j] = acc[i][j] + bh;
        ss[j] = sigmoidf_(acc[i][4 + j] + bg);
        o[j] = hh[j] * ss[j];
      }
      const size_t base = (size_t)m * p.ldc + j0;
      if (j0 + 3 < p.O && p.c_vec) {
        *reinterpret_cast<float4*>(p.out0 + base) = make_float4(o[0], o[1], o[2], o[3]);
        if (p.out1) *reinterpret_cast<float4*>(p.out1 + base) = make_float4(hh[0], hh[1], hh[2], hh[3]);
        if (p.out2) *reinterpret_cast<float4*>(p.out2 + base) = make_float4(ss[0], ss[1], ss[2], ss[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j0 + j < p.O) {
            p.out0[base + j] = o[j];
            if (p.out1) p.out1[base + j] = hh[j];
            if (p.out2) p.out2[base + j] = ss[j];
          }
      }
    } else {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c0 = n0 + 64 * q + tx * 4;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v = acc[i][4 * q + j];
          if (EPI == EPI_BIAS_ACT) {
            const int col = c0 + j;
            if (p.bias0 && col < p.N) v += p.bias0[col];
            if (p.act == EXVAE_ACT_SIGMOID) v = sigmoidf_(v);
            else if (p.act == EXVAE_ACT_HARDTANH) v = fminf(fmaxf(v, p.lo), p.hi);
            else if (p.act == EXVAE_ACT_RELU) v = fmaxf(v, 0.f);
          }
          o[j] = v;
        }
        float* dst = (EPI == EPI_SPLITK) ? p.out0 + ((size_t)blockIdx.z * p.M + m) * p.ldc : p.out0 + (size_t)m * p.ldc;
        if (c0 + 3 < p.N && p.c_vec) {
          *reinterpret_cast<float4*>(dst + c0) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c0 + j < p.N) dst[c0 + j] = o[j];
        }
      }
    }
  }
}

// out rows m < mseg -> out0[m], else out1[m - mseg]
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, int S, int M, int N, int mseg,
                                                            float* __restrict__ out0, float* __restrict__ out1,
                                                            int accumulate) {
  const size_t total = (size_t)M * N;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int s = 0; s < S; ++s) a += part[(size_t)s * total + e];
    const int m = (int)(e / N);
    float* dst = m < mseg ? out0 + e : out1 + (e - (size_t)mseg * N);
    *dst = accumulate ? *dst + a : a;
  }
}

// stage 1 of the column sum (bias gradients): block = 32 columns x 8 row lanes
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ src, int R, int ncols, int rows_per,
                                                             float* __restrict__ part) {
  __shared__ float sh[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  const int r0 = blockIdx.y * rows_per, r1 = min(R, r0 + rows_per);
  float a = 0.f;
  if (col < ncols)
    for (int r = r0 + ry; r < r1; r += 8) a += src[(size_t)r * ncols + col];
  sh[ry][cx] = a;
  __syncthreads();
  if (ry == 0 && col < ncols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][cx];
    part[(size_t)blockIdx.y * ncols + col] = t;
  }
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ part, int S, int ncols, int seg,
                                                           float* __restrict__ out0, float* __restrict__ out1,
                                                           int accumulate) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  float a = 0.f;
  for (int s = 0; s < S; ++s) a += part[(size_t)s * ncols + col];
  float* dst = col < seg ? (out0 ? out0 + col : nullptr) : (out1 ? out1 + (col - seg) : nullptr);
  if (dst) *dst = accumulate ? *dst + a : a;
}

// dcat[r, j] = dout*sig ; dcat[r, O+j] = dout*h*sig*(1-sig)      (d/dh and d/dg of h*sigmoid(g))
__global__ void __launch_bounds__(256) gated_dpre_kernel(const float* __restrict__ dout, const float* __restrict__ h,
                                                         const float* __restrict__ sig, long long R, int O,
                                                         float* __restrict__ dcat) {
  const long long total = R * O;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / O;
    const int j = (int)(e - r * O);
    const float d = dout[e], s = sig[e], ov = h[e];   // ov = layer output h*s
    dcat[r * 2 * O + j] = d * s;
    dcat[r * 2 * O + O + j] = d * ov * (1.f - s);
  }
}
// dpre = dout * act'(out)
__global__ void __launch_bounds__(256) act_dpre_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                       long long n, int act, float lo, float hi,
                                                       float* __restrict__ dpre) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float o = out[e];
    float d = dout[e];
    if (act == EXVAE_ACT_SIGMOID) d *= o * (1.f - o);
    else if (act == EXVAE_ACT_HARDTANH) d = (o > lo && o < hi) ? d : 0.f;
    else if (act == EXVAE_ACT_RELU) d = o > 0.f ? d : 0.f;
    dpre[e] = d;
  }
}

// Tensor-core backward staging, ONE pass: pre-activation gradient dcat (the fp32 GEMM operand; the GEMM splits it
// into tf32 hi/lo parts in shared memory) + per-block column sums (bias gradients).  Replaces dpre + column-sum
// (2 kernels, 1 extra round trip through HBM).
//   MODE 0: gated   dcat[r, j] = dout*sig ; dcat[r, O+j] = dout*out*(1-sig), out = h*sig   (ncat = 2*O)
//   MODE 1: linear  dcat[r, j] = dout * act'(out)                                      (ncat = O)
// Thread mapping: a thread owns ONE group of VW columns (VW = 4: 16-byte accesses when O % 4 == 0, else VW = 1) and one
// of RL = 256 / (O/VW) row lanes, so narrow layers (the 40-wide latent heads) and the 300-wide trunk both keep most of
// the 256 threads busy.  A block covers RL*iters rows; the RL lanes' column sums are combined through shared memory,
// so cs_part holds one row per block.  Requires O / VW <= 256.
template <int VW>
struct VecT { typedef float4 type; };
template <>
struct VecT<1> { typedef float type; };
template <int VW>
__device__ __forceinline__ void vload(const float* p, float (&v)[VW]) {
  const typename VecT<VW>::type t = *reinterpret_cast<const typename VecT<VW>::type*>(p);
  memcpy(v, &t, sizeof(t));
}
template <int VW>
__device__ __forceinline__ void vstore(float* p, const float (&v)[VW]) {
  typename VecT<VW>::type t;
  memcpy(&t, v, sizeof(t));
  *reinterpret_cast<typename VecT<VW>::type*>(p) = t;
}
template <int MODE, int VW>
__global__ void __launch_bounds__(256) dpre_colsum_kernel(const float* __restrict__ dout,
                                                          const float* __restrict__ h,
                                                          const float* __restrict__ sig_or_out, int R, int O,
                                                          int act, float lo, float hi, int RL, int iters,
                                                          float* __restrict__ dsplit,
                                                          float* __restrict__ cs_part, int ldd) {
  extern __shared__ float sh_cs[];               // [RL][ncat]
  const int ncat = MODE == 0 ? 2 * O : O;
  if (ldd == 0) ldd = ncat;                      // row pitch of dsplit (conv layers with ncat % 4 != 0 pad it to 4)
  const int Ov = O / VW;
  const int rbase = blockIdx.x * RL * iters;
  const int cv = threadIdx.x % Ov, lanei = threadIdx.x / Ov;
  if (lanei < RL) {
    float a0[VW], a1[VW];
#pragma unroll
    for (int k = 0; k < VW; ++k) a0[k] = a1[k] = 0.f;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
      const int r = rbase + i * RL + lanei;
      if (r >= R) break;
      const size_t e = (size_t)r * O + VW * cv;
      float d[VW];
      vload<VW>(dout + e, d);
      if (MODE == 0) {
        float sg[VW], ov[VW], dh[VW], dg[VW];
        vload<VW>(sig_or_out + e, sg);
        vload<VW>(h + e, ov);                    // layer output h*s
#pragma unroll
        for (int k = 0; k < VW; ++k) {
          dh[k] = d[k] * sg[k];
          dg[k] = d[k] * ov[k] * (1.f - sg[k]);
          a0[k] += dh[k];
          a1[k] += dg[k];
        }
        vstore<VW>(dsplit + (size_t)r * ldd + VW * cv, dh);
        vstore<VW>(dsplit + (size_t)r * ldd + O + VW * cv, dg);
      } else {
        if (act != EXVAE_ACT_NONE) {
          float o[VW];
          vload<VW>(sig_or_out + e, o);
#pragma unroll
          for (int k = 0; k < VW; ++k) {
            if (act == EXVAE_ACT_SIGMOID) d[k] *= o[k] * (1.f - o[k]);
            else if (act == EXVAE_ACT_HARDTANH) d[k] = (o[k] > lo && o[k] < hi) ? d[k] : 0.f;
            else if (act == EXVAE_ACT_RELU) d[k] = o[k] > 0.f ? d[k] : 0.f;
          }
        }
        vstore<VW>(dsplit + (size_t)r * ldd + VW * cv, d);
#pragma unroll
        for (int k = 0; k < VW; ++k) a0[k] += d[k];
      }
    }
#pragma unroll
    for (int k = 0; k < VW; ++k) {
      sh_cs[lanei * ncat + VW * cv + k] = a0[k];
      if (MODE == 0) sh_cs[lanei * ncat + O + VW * cv + k] = a1[k];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < ncat; c += 256) {
    float t = 0.f;
    for (int l = 0; l < RL; ++l) t += sh_cs[l * ncat + c];
    cs_part[(size_t)blockIdx.x * ncat + c] = t;
  }
}

// ONE launch that finishes a layer's parameter gradients: blocks [0, nred) sum the split-K partials of dW
// (part [S][M][N] -> out0 rows < mseg, out1 the rest), the following ceil(ncols/32) blocks reduce the staging
// kernel's column sums cs [S2][ncols] into the bias gradients (32 columns x 8 row lanes per block).
__global__ void __launch_bounds__(256) dw_finish_kernel(const float* __restrict__ part, int S, int M, int N, int mseg,
                                                        float* __restrict__ out0, float* __restrict__ out1, int nred,
                                                        const float* __restrict__ cs, int S2, int ncols,
                                                        float* __restrict__ db0, float* __restrict__ db1,
                                                        int accumulate) {
  if ((int)blockIdx.x < nred) {
    const size_t total = (size_t)M * N;
    // 16-byte path (N % 4 == 0 and aligned planes / destinations: a float4 never straddles a row or the out0 | out1
    // boundary): a quarter of the threads and instructions of the scalar loop -- this pass is launch / latency bound and,
    // for the first layer, sits between the last GEMM of the backward and the optimizer
    const bool v4 = (N & 3) == 0 && ((reinterpret_cast<uintptr_t>(part) | reinterpret_cast<uintptr_t>(out0) |
                                      reinterpret_cast<uintptr_t>(out1)) & 15) == 0;
    if (v4) {
      const size_t total4 = total >> 2;
      for (size_t e4 = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e4 < total4; e4 += (size_t)nred * blockDim.x) {
        const size_t e = e4 << 2;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < S; ++s) {
          const float4 t = *reinterpret_cast<const float4*>(part + (size_t)s * total + e);
          a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        const int m = (int)(e / N);
        float4* dst = reinterpret_cast<float4*>(m < mseg ? out0 + e : out1 + (e - (size_t)mseg * N));
        if (accumulate) {
          const float4 o = *dst;
          a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
        }
        *dst = a;
      }
      return;
    }
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)nred * blockDim.x) {
      float a = 0.f;
      for (int s = 0; s < S; ++s) a += part[(size_t)s * total + e];
      const int m = (int)(e / N);
      float* dst = m < mseg ? out0 + e : out1 + (e - (size_t)mseg * N);
      *dst = accumulate ? *dst + a : a;
    }
    return;
  }
  __shared__ float sh[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = ((int)blockIdx.x - nred) * 32 + cx;
  float a = 0.f;
  if (col < ncols)
    for (int r = ry; r < S2; r += 8) a += cs[(size_t)r * ncols + col];
  sh[ry][cx] = a;
  __syncthreads();
  if (ry == 0 && col < ncols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][cx];
    float* dst = col < mseg ? (db0 ? db0 + col : nullptr) : (db1 ? db1 + (col - mseg) : nullptr);
    if (dst) *dst = accumulate ? *dst + t : t;
  }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline int ew_blocks(long long n) { return (int)std::min<long long>((n + 255) / 256, 148LL * 16); }

template <int AL, int BL, int EPI>
int launch_gemm(const GemmP& p, int splits, cudaStream_t st) {
  dim3 grid(EPI == EPI_GATED ? ceil_div(p.O, GN / 2) : ceil_div(p.N, GN), ceil_div(p.M, GM), splits);
  sgemm_kernel<AL, BL, EPI><<<grid, 256, 0, st>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? EXVAE_OK : (int)e;
}

struct BwdPlan {
  int S, kchunk, S2, rows_per;
  size_t off_dcat, off_part, off_cs, bytes;
};
// ncat = number of pre-activation columns (2*O gated, O linear); need_dcat: a dpre buffer is needed
inline BwdPlan bwd_plan(int R, int K, int ncat, bool need_dcat) {
  BwdPlan b;
  const int tiles = ceil_div(ncat, GM) * ceil_div(K, GN);
  // split the R-reduction so that tiles*S fills ONE wave of the 2-CTA/SM residency (never 1.x waves)
  const int slots = 2 * sm_count();
  int S = std::max(1, slots / tiles);
  S = std::max(1, std::min(S, ceil_div(R, 4 * GK)));
  b.kchunk = ceil_div(ceil_div(R, S), GK) * GK;
  b.S = ceil_div(R, b.kchunk);
  b.S2 = std::max(1, std::min(64, ceil_div(R, 256)));
  b.rows_per = ceil_div(R, b.S2);
  size_t off = 0;
  b.off_dcat = off;
  if (need_dcat) off += align_up(sizeof(float) * (size_t)R * ncat, 256);
  b.off_part = off;
  off += align_up(sizeof(float) * (size_t)b.S * ncat * K, 256);
  b.off_cs = off;
  off += align_up(sizeof(float) * (size_t)b.S2 * ncat, 256);
  b.bytes = off;
  return b;
}

// shared tail of both backward passes: dx, dW (split-K over rows + reduce), db
int dense_bwd_common(const float* x, const float* W0, const float* W1, const float* dcat, int R, int K, int ncat, int oseg,
                     float* dx, float* dW0, float* dW1, float* db0, float* db1, const BwdPlan& plan, char* ws,
                     int accumulate, cudaStream_t st) {
  int rc;
  if (dx) {  // dx[R,K] = dcat[R,ncat] . Wcat[ncat,K]
    GemmP p{};
    p.A = dcat; p.lda = ncat; p.B0 = W0; p.B1 = W1 ? W1 : W0; p.ldb = K; p.bseg = W1 ? oseg : ncat;
    p.M = R; p.N = K; p.K = ncat; p.kchunk = ncat; p.out0 = dx; p.ldc = K;
    p.c_vec = (K % 4 == 0) && al16(dx);
    p.a_vec = (ncat % 4 == 0) && al16(dcat);
    p.b_vec = (K % 4 == 0) && al16(W0) && (!W1 || al16(W1));
    rc = launch_gemm<L_KC, L_XC, EPI_PLAIN>(p, 1, st);
    if (rc) return rc;
  }
  {  // dWcat[ncat,K] = dcat^T[ncat,R] . x[R,K]   (reduction over the R rows, split across CTAs)
    float* part = reinterpret_cast<float*>(ws + plan.off_part);
    GemmP p{};
    p.A = dcat; p.lda = ncat; p.B0 = x; p.B1 = x; p.ldb = K; p.bseg = R;
    p.M = ncat; p.N = K; p.K = R; p.kchunk = plan.kchunk; p.out0 = part; p.ldc = K;
    p.c_vec = (K % 4 == 0) && (((size_t)ncat * K) % 4 == 0) && al16(part);
    p.a_vec = (ncat % 4 == 0) && al16(dcat);
    p.b_vec = (K % 4 == 0) && al16(x);
    rc = launch_gemm<L_XC, L_XC, EPI_SPLITK>(p, plan.S, st);
    if (rc) return rc;
    splitk_reduce_kernel<<<ew_blocks((long long)ncat * K), 256, 0, st>>>(part, plan.S, ncat, K, oseg, dW0,
                                                                         dW1 ? dW1 : dW0, accumulate);
    EXVAE_CUDA(cudaGetLastError());
  }
  if (db0 || db1) {
    float* cs = reinterpret_cast<float*>(ws + plan.off_cs);
    dim3 g1(ceil_div(ncat, 32), plan.S2);
    colsum_partial_kernel<<<g1, 256, 0, st>>>(dcat, R, ncat, plan.rows_per, cs);
    EXVAE_CUDA(cudaGetLastError());
    colsum_final_kernel<<<ceil_div(ncat, 256), 256, 0, st>>>(cs, plan.S2, ncat, oseg, db0, db1, accumulate);
    EXVAE_CUDA(cudaGetLastError());
  }
  return EXVAE_OK;
}

// ------------------------------------------------------------------ tensor-core (3xTF32) plumbing
// forward workspace: gated layers keep [Wh ; Wg] concatenated as one [2*O, K] operand (reused by the backward);
// plain linear layers need none (the GEMM reads x and W where they lie).
struct FwdWs {
  size_t off_w, bytes;
};
inline FwdWs fwd_ws_layout(int R, int K, int OC, bool gated) {
  (void)R;
  FwdWs f;
  f.off_w = 0;
  f.bytes = gated ? align_up(sizeof(float) * (size_t)OC * K, 256) : 0;
  return f;
}
inline bool tc_ok(int R, int K, int OC, const void* x) {
  return tc_enabled() && tc_dims_ok(K) && tc_dims_ok(OC) && OC <= 2048 && al16(x) && R > 0;
}
struct TcBwdPlan {
  int S, kchunk, S2, RL, iters;
  size_t off_dcat, off_dsplit, off_w, off_part, off_cs, bytes;
};
// upper bound of the staging kernel's block count over all geometries (RL*iters >= 1 row per block, and at most
// ~592 blocks unless 16 rows per block are not enough)
inline int stage_blocks_max(int R) { return std::max(ceil_div(R, 16), std::min(R, 600)); }
inline TcBwdPlan tc_bwd_plan(int R, int K, int ncat) {
  TcBwdPlan b;
  // Split the reduction over the R rows so that the persistent kernel's tiles fill whole waves of SMs: cost (in
  // k-blocks of 32 rows) = waves * k-blocks per split + the write/read of one more partial [ncat,K] per split
  // (~3 TB/s against ~0.77 us per k-block).  Accuracy: the TMEM accumulator is fp32 with truncating adds, so one
  // accumulation chain is capped at 2560 rows; the fp32 split-K reduction (round-to-nearest) combines the chunks.
  const int tiles = ceil_div(ncat, 128) * ceil_div(K, 128);
  const int sms = sm_count();
  const double per_split = 2.0 * ncat * (double)K * 4.0 / 3.0e12 / 0.77e-6;
  const int s_lo = std::max(1, ceil_div(R, 2560)), s_hi = std::max(s_lo, std::min(ceil_div(R, 128), 96));
  int S = s_lo;
  double best = 1e30;
  for (int c = s_lo; c <= s_hi; ++c) {
    const int kc = ceil_div(ceil_div(R, c), 32) * 32;
    const int sp = ceil_div(R, kc);                      // splits that actually hold rows
    const double cost = (double)ceil_div(tiles * sp, sms) * (kc / 32) + per_split * sp;
    if (cost < best) { best = cost; S = c; }
  }
  b.kchunk = ceil_div(ceil_div(R, S), 32) * 32;
  b.S = ceil_div(R, b.kchunk);
  b.S2 = 0; b.RL = 1; b.iters = 1;       // staging kernel geometry: filled by tc_stage_geometry (depends on the layer's O)
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  b.off_dcat = 0;
  b.off_dsplit = take(sizeof(float) * (size_t)R * ncat);
  b.off_w = take(sizeof(float) * (size_t)ncat * K);
  b.off_part = take(sizeof(float) * (size_t)b.S * ncat * K);
  b.off_cs = take(sizeof(float) * (size_t)stage_blocks_max(R) * ncat);
  b.bytes = off;
  return b;
}
// geometry of dpre_colsum_kernel for a layer with O pre-activation columns per segment
inline void tc_stage_geometry(TcBwdPlan& b, int R, int O) {
  b.RL = std::max(1, 256 / (O % 4 == 0 ? O / 4 : O));
  b.iters = std::min(16, std::max(1, ceil_div(R, 592 * b.RL)));
  b.S2 = ceil_div(R, b.RL * b.iters);
}

// one non-blocking side stream + fork/join events per device, created on first use (i.e. in an eager warm-up step,
// before any graph capture)
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool pending = false;      // a deferred dW finish has been queued and not yet joined (exvae_dense_bwd_flush)
};
// exvae_dense_bwd_defer_finish: the dW finish of the LARGE layers is forked and NOT joined by the backward call
bool g_defer_finish = false;
inline SideStream* side_stream() {
  static SideStream table[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  SideStream& s = table[dev];
  if (!s.stream) {
    // highest priority: the forked dW GEMMs belong to the main chain of the step (the exemplar-prior branch runs its
    // whole-GPU kernels on a default = low priority stream next to them)
    int prio_lo = 0, prio_hi = 0;
    if (cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess) { (void)cudaGetLastError(); prio_hi = 0; }
    if (cudaStreamCreateWithPriority(&s.stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) {
      (void)cudaGetLastError();
      s.stream = nullptr;
      return nullptr;          // fall back to the serial order
    }
  }
  return &s;
}

// dx / dW / db on the tensor cores.  The pre-activation gradient arrives already staged by dpre_colsum_kernel:
// dcat in ws+off_dsplit, column-sum partials in ws+off_cs.  wcat_in: the forward's [W0 ; W1] copy (gated) or null.
int dense_bwd_tc(const float* x, const float* W0, const float* W1, int R, int K, int ncat, int oseg,
                 float* dx, float* dW0, float* dW1, float* db0, float* db1, const float* wcat_in,
                 const TcBwdPlan& plan, char* ws, int accumulate, cudaStream_t st) {
  float* dsplit = reinterpret_cast<float*>(ws + plan.off_dsplit);
  int rc;
  const float* wsp = W1 ? wcat_in : W0;
  if (dx && !wsp) {
    float* w_own = reinterpret_cast<float*>(ws + plan.off_w);
    rc = tc_concat2(W0, W1, (size_t)oseg * K, w_own, st);
    if (rc) return rc;
    wsp = w_own;
  }
  // Small layers (the decoder and the batch-only heads: a few tiles each) leave most SMs idle, and dx and dW only share
  // their inputs: fork dW + the parameter-gradient finish onto a side stream and join afterwards (plain event
  // fork/join, so a capturing stream turns it into two parallel graph branches).
  SideStream* side = (dx && R <= 4096) ? side_stream() : nullptr;
  cudaStream_t sw = st;
  if (side) {
    EXVAE_CUDA(cudaEventRecord(side->fork, st));
    EXVAE_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    sw = side->stream;
  }
  if (dx) {  // dx[R,K] = dcat[R,ncat] . Wcat[ncat,K]  : A K-major, B MN-major ([ncat rows][K cols])
    TcGemm g{};
    g.a = dsplit; g.a_rows = R; g.a_cols = ncat; g.a_mn = false;
    g.b = wsp; g.b_rows = ncat; g.b_cols = K; g.b_mn = true;
    g.M = R; g.N = K; g.K = ncat; g.epi = TC_PLAIN; g.out0 = dx; g.ldc = K;
    rc = tc_gemm_launch(g, st);
    if (rc) return rc;
  }
  {  // dWcat[ncat,K] = dcat^T . x : A MN-major ([R rows][ncat cols]), B MN-major ([R rows][K cols])
    float* part = reinterpret_cast<float*>(ws + plan.off_part);
    TcGemm g{};
    g.a = dsplit; g.a_rows = R; g.a_cols = ncat; g.a_mn = true;
    g.b = x; g.b_rows = R; g.b_cols = K; g.b_mn = true;
    g.M = ncat; g.N = K; g.K = R; g.epi = TC_SPLITK; g.out0 = part; g.ldc = K;
    g.splits = plan.S; g.kchunk = plan.kchunk;
    rc = tc_gemm_launch(g, sw);
    if (rc) return rc;
    // split-K reduction of dW and the bias gradients (column sums of the staging blocks) in one launch.
    // Large layers with a deferred finish: this launch-bound pass (5-14 us) goes to the side stream and is joined by
    // exvae_dense_bwd_flush (or by the next fork/join on that stream), so that it runs next to the following layer's
    // staging kernel instead of in front of it.  The caller keeps `ws` alive until the flush.
    SideStream* dside = (!side && g_defer_finish && accumulate) ? side_stream() : nullptr;
    if (dside) {
      EXVAE_CUDA(cudaEventRecord(dside->fork, st));
      EXVAE_CUDA(cudaStreamWaitEvent(dside->stream, dside->fork, 0));
      sw = dside->stream;
      dside->pending = true;
    }
    // (same test as the kernel's 16-byte path: then a quarter of the blocks)
    const bool v4 = (K & 3) == 0 && al16(part) && al16(dW0) && al16(dW1 ? dW1 : dW0);
    const int nred = ew_blocks((long long)ncat * K / (v4 ? 4 : 1));
    const int ncs = (db0 || db1) ? ceil_div(ncat, 32) : 0;
    dw_finish_kernel<<<nred + ncs, 256, 0, sw>>>(part, plan.S, ncat, K, oseg, dW0, dW1 ? dW1 : dW0, nred,
                                                 reinterpret_cast<const float*>(ws + plan.off_cs), plan.S2, ncat, db0,
                                                 db1, accumulate);
    EXVAE_CUDA(cudaGetLastError());
  }
  if (side) {
    EXVAE_CUDA(cudaEventRecord(side->join, side->stream));
    EXVAE_CUDA(cudaStreamWaitEvent(st, side->join, 0));
    side->pending = false;       // the side stream is in order: this join also covers an earlier deferred finish
  }
  return EXVAE_OK;
}

}  // namespace
}  // namespace exvae

using namespace exvae;

extern "C" size_t exvae_dense_fwd_workspace_bytes(int R, int K, int O, int gated) {
  if (R <= 0 || K <= 0 || O <= 0) return 0;
  return fwd_ws_layout(R, K, gated ? 2 * O : O, gated != 0).bytes;
}

extern "C" int exvae_gated_dense_fwd(const float* x, const float* Wh, const float* bh, const float* Wg, const float* bg,
                                     int R, int K, int O, float* out, float* sig, void* ws,
                                     size_t ws_bytes, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(x && Wh && Wg && out && R > 0 && K > 0 && O > 0);
  cudaStream_t st = as_stream(stream);
  const FwdWs f = fwd_ws_layout(R, K, 2 * O, true);
  if (ws && ws_bytes >= f.bytes && tc_ok(R, K, 2 * O, x) && al16(Wh) && al16(Wg) && al16(ws)) {
    // [Wh ; Wg] as ONE operand: when the caller keeps the two weights adjacent in memory (layers.GatedDense packs them
    // into one [2*O, K] buffer) Wh already IS that operand and the copy disappears
    const float* wsp = Wh;
    if (Wg != Wh + (size_t)O * K) {
      float* wown = reinterpret_cast<float*>(static_cast<char*>(ws) + f.off_w);
      int rc = tc_concat2(Wh, Wg, (size_t)O * K, wown, st);
      if (rc) return rc;
      wsp = wown;
    }
    TcGemm g{};
    g.a = x; g.a_rows = R; g.a_cols = K; g.a_mn = false;
    g.b = wsp; g.b_rows = 2 * O; g.b_cols = K; g.b_mn = false;
    g.M = R; g.N = O; g.K = K; g.epi = TC_GATED; g.gated_O = O;
    g.bias0 = bh; g.bias1 = bg; g.out0 = out; g.out1 = nullptr; g.out2 = sig; g.ldc = O;
    return tc_gemm_launch(g, st);
  }
  GemmP p{};
  p.A = x; p.lda = K; p.B0 = Wh; p.B1 = Wg; p.ldb = K; p.M = R; p.N = 2 * O; p.K = K; p.kchunk = K;
  p.bias0 = bh; p.bias1 = bg; p.out0 = out; p.out1 = nullptr; p.out2 = sig; p.ldc = O; p.O = O;
  p.a_vec = (K % 4 == 0) && al16(x);
  p.b_vec = (K % 4 == 0) && al16(Wh) && al16(Wg);
  p.c_vec = (O % 4 == 0) && al16(out) && (!sig || al16(sig));
  return launch_gemm<L_KC, L_KC, EPI_GATED>(p, 1, st);
}

extern "C" size_t exvae_gated_dense_bwd_workspace_bytes(int R, int K, int O) {
  if (R <= 0 || K <= 0 || O <= 0) return 0;
  return std::max(bwd_plan(R, K, 2 * O, true).bytes, tc_bwd_plan(R, K, 2 * O).bytes);
}

extern "C" int exvae_gated_dense_bwd(const float* x, const float* Wh, const float* Wg, const float* out,
                                     const float* sig, const float* dout, int R, int K, int O, float* dx, float* dWh,
                                     float* dbh, float* dWg, float* dbg, const void* fwd_ws, size_t fwd_ws_bytes,
                                     void* ws, size_t ws_bytes, int accumulate, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(x && Wh && Wg && out && sig && dout && dWh && dWg && ws && R > 0 && K > 0 && O > 0);
  cudaStream_t st = as_stream(stream);
  char* w = static_cast<char*>(ws);
  const bool tc = tc_ok(R, K, 2 * O, x) && (O % 4 == 0 || O <= 256) && al16(Wh) && al16(Wg) && al16(ws) &&
                  al16(dout) && al16(out) && al16(sig);
  if (tc) {
    TcBwdPlan plan = tc_bwd_plan(R, K, 2 * O);
    if (ws_bytes < plan.bytes) return EXVAE_ERR_WORKSPACE;
    tc_stage_geometry(plan, R, O);
    const size_t sh = sizeof(float) * plan.RL * 2 * O;
    auto stage_kern = (O % 4 == 0) ? dpre_colsum_kernel<0, 4> : dpre_colsum_kernel<0, 1>;
    stage_kern<<<plan.S2, 256, sh, st>>>(dout, out, sig, R, O, 0, 0.f, 0.f, plan.RL, plan.iters,
                                                         reinterpret_cast<float*>(w + plan.off_dsplit),
                                                         reinterpret_cast<float*>(w + plan.off_cs), 0);
    EXVAE_CUDA(cudaGetLastError());
    const FwdWs f = fwd_ws_layout(R, K, 2 * O, true);
    const bool reuse = fwd_ws && fwd_ws_bytes >= f.bytes && al16(fwd_ws);
    const float* wsp = (Wg == Wh + (size_t)O * K) ? Wh          // adjacent weights: no [Wh ; Wg] copy exists or is needed
                       : reuse ? reinterpret_cast<const float*>(static_cast<const char*>(fwd_ws) + f.off_w) : nullptr;
    return dense_bwd_tc(x, Wh, Wg, R, K, 2 * O, O, dx, dWh, dWg, dbh, dbg, wsp, plan, w, accumulate, st);
  }
  const BwdPlan plan = bwd_plan(R, K, 2 * O, true);
  if (ws_bytes < plan.bytes) return EXVAE_ERR_WORKSPACE;
  float* dcat = reinterpret_cast<float*>(w + plan.off_dcat);
  gated_dpre_kernel<<<ew_blocks((long long)R * O), 256, 0, st>>>(dout, out, sig, R, O, dcat);
  EXVAE_CUDA(cudaGetLastError());
  return dense_bwd_common(x, Wh, Wg, dcat, R, K, 2 * O, O, dx, dWh, dWg, dbh, dbg, plan, w, accumulate, st);
}

extern "C" int exvae_linear_fwd(const float* x, const float* W, const float* b, int R, int K, int O, int act, float lo,
                                float hi, float* out, void* ws, size_t ws_bytes, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(x && W && out && R > 0 && K > 0 && O > 0);
  EXVAE_CHECK_ARG(act >= EXVAE_ACT_NONE && act <= EXVAE_ACT_RELU);
  cudaStream_t st = as_stream(stream);
  (void)ws; (void)ws_bytes;
  if (tc_ok(R, K, O, x) && al16(W)) {
    TcGemm g{};
    g.a = x; g.a_rows = R; g.a_cols = K; g.a_mn = false;
    g.b = W; g.b_rows = O; g.b_cols = K; g.b_mn = false;
    g.M = R; g.N = O; g.K = K; g.epi = TC_BIAS_ACT;
    g.bias0 = b; g.out0 = out; g.ldc = O; g.act = act; g.lo = lo; g.hi = hi;
    return tc_gemm_launch(g, st);
  }
  GemmP p{};
  p.A = x; p.lda = K; p.B0 = W; p.B1 = W; p.ldb = K; p.M = R; p.N = O; p.K = K; p.kchunk = K;
  p.bias0 = b; p.out0 = out; p.ldc = O; p.act = act; p.lo = lo; p.hi = hi;
  p.a_vec = (K % 4 == 0) && al16(x);
  p.b_vec = (K % 4 == 0) && al16(W);
  p.c_vec = (O % 4 == 0) && al16(out);
  return launch_gemm<L_KC, L_KC, EPI_BIAS_ACT>(p, 1, st);
}

extern "C" size_t exvae_linear_bwd_workspace_bytes(int R, int K, int O) {
  if (R <= 0 || K <= 0 || O <= 0) return 0;
  return std::max(bwd_plan(R, K, O, true).bytes, tc_bwd_plan(R, K, O).bytes);
}

extern "C" int exvae_linear_bwd(const float* x, const float* W, const float* out, const float* dout, int R, int K, int O,
                                int act, float lo, float hi, float* dx, float* dW, float* db, const void* fwd_ws,
                                size_t fwd_ws_bytes, void* ws, size_t ws_bytes, int accumulate,
                                exvae_stream_t stream) {
  EXVAE_CHECK_ARG(x && W && dout && dW && ws && R > 0 && K > 0 && O > 0);
  EXVAE_CHECK_ARG(act == EXVAE_ACT_NONE || out != nullptr);
  cudaStream_t st = as_stream(stream);
  char* w = static_cast<char*>(ws);
  const bool tc = tc_ok(R, K, O, x) && O <= 1024 && al16(W) && al16(ws) && al16(dout) && (act == EXVAE_ACT_NONE || al16(out));
  if (tc) {
    TcBwdPlan plan = tc_bwd_plan(R, K, O);
    if (ws_bytes < plan.bytes) return EXVAE_ERR_WORKSPACE;
    tc_stage_geometry(plan, R, O);
    const size_t sh = sizeof(float) * plan.RL * O;
    auto stage_kern = (O % 4 == 0) ? dpre_colsum_kernel<1, 4> : dpre_colsum_kernel<1, 1>;
    stage_kern<<<plan.S2, 256, sh, st>>>(dout, nullptr, out, R, O, act, lo, hi, plan.RL, plan.iters,
                                                         reinterpret_cast<float*>(w + plan.off_dsplit),
                                                         reinterpret_cast<float*>(w + plan.off_cs), 0);
    EXVAE_CUDA(cudaGetLastError());
    (void)fwd_ws; (void)fwd_ws_bytes;
    return dense_bwd_tc(x, W, nullptr, R, K, O, O, dx, dW, nullptr, db, nullptr, nullptr, plan, w, accumulate, st);
  }
  const BwdPlan plan = bwd_plan(R, K, O, true);
  if (ws_bytes < plan.bytes) return EXVAE_ERR_WORKSPACE;
  const float* dpre = dout;
  if (act != EXVAE_ACT_NONE) {
    float* buf = reinterpret_cast<float*>(w + plan.off_dcat);
    act_dpre_kernel<<<ew_blocks((long long)R * O), 256, 0, st>>>(dout, out, (long long)R * O, act, lo, hi, buf);
    EXVAE_CUDA(cudaGetLastError());
    dpre = buf;
  }
  return dense_bwd_common(x, W, nullptr, dpre, R, K, O, O, dx, dW, nullptr, db, nullptr, plan, w, accumulate, st);
}

extern "C" int exvae_gemm_backend(void) { return tc_enabled() ? 1 : 0; }
extern "C" int exvae_dense_bwd_defer_finish(int on) {
  const int prev = g_defer_finish ? 1 : 0;
  g_defer_finish = on != 0;
  return prev;
}

extern "C" int exvae_dense_bwd_flush(exvae_stream_t stream) {
  SideStream* side = side_stream();
  if (side && side->pending) {
    EXVAE_CUDA(cudaEventRecord(side->join, side->stream));
    EXVAE_CUDA(cudaStreamWaitEvent(as_stream(stream), side->join, 0));
    side->pending = false;
  }
  return EXVAE_OK;
}

extern "C" int exvae_gemm_set_trace(uint64_t* buf) {
  tc_set_trace(reinterpret_cast<unsigned long long*>(buf));
  return EXVAE_OK;
}


// =====================================================================================================================
// K4 — convolution entry points (GatedConv2d / Conv2d, utils/nn.py:72-114; weight-normed convs of models/fully_conv.py)
// =====================================================================================================================
namespace exvae {
namespace {

struct ConvPlan {
  int OH, OW, taps;
  long long R;
  int implicit, cpad, Kp;          // forward operand: [ncat][Kp], K index = tap*cpad + c
  int Kpc;                         // materialised patch rows: pitch round4(taps*Cin), channel pitch Cin
  int dx_implicit, cpad_dx, Kp_dx; // stride-1 input gradient as a convolution of dcat: operand [Cin][taps*cpad_dx]
  int ldd;                         // row pitch of dcat (ncat rounded up to 4)
  int bh, bn, bh_dx, bn_dx;
};
inline ConvPlan conv_plan(int N, int H, int W, int Cin, int KH, int KW, int stride, int pad, int ncat) {
  ConvPlan c{};
  c.OH = (H + 2 * pad - KH) / stride + 1;
  c.OW = (W + 2 * pad - KW) / stride + 1;
  c.taps = KH * KW;
  c.R = (long long)N * c.OH * c.OW;
  c.Kpc = ceil_div(c.taps * Cin, 4) * 4;
  c.ldd = ceil_div(ncat, 4) * 4;
  const bool tc = tc_enabled();
  c.implicit = tc && Cin >= 16 && Cin % 4 == 0 && tc_conv_tiling(c.OH, c.OW, &c.bh, &c.bn);
  c.cpad = c.implicit ? ceil_div(Cin, 32) * 32 : Cin;
  c.Kp = c.implicit ? c.taps * c.cpad : c.Kpc;
  c.dx_implicit = tc && stride == 1 && ncat >= 16 && ncat % 4 == 0 && KH - 1 - pad >= 0 && KW - 1 - pad >= 0 &&
                  tc_conv_tiling(H, W, &c.bh_dx, &c.bn_dx);
  c.cpad_dx = ceil_div(ncat, 32) * 32;
  c.Kp_dx = c.taps * c.cpad_dx;
  return c;
}

struct ConvBwdWs {
  size_t off_dcat, off_cs, off_col, off_part, bytes;
  TcBwdPlan gp;
};
inline ConvBwdWs conv_bwd_ws(const ConvPlan& c, int ncat, int O) {
  ConvBwdWs w;
  w.gp = tc_bwd_plan((int)c.R, c.Kpc, ncat);
  tc_stage_geometry(w.gp, (int)c.R, O);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  w.off_dcat = take(sizeof(float) * (size_t)c.R * c.ldd);
  w.off_cs = take(sizeof(float) * (size_t)stage_blocks_max((int)c.R) * ncat);
  w.off_col = take(sizeof(float) * (size_t)c.R * c.Kpc);
  w.off_part = take(sizeof(float) * (size_t)w.gp.S * ncat * c.Kpc);
  w.bytes = off;
  return w;
}

}  // namespace
}  // namespace exvae

/* plan[0..8] = {OH, OW, implicit, cpad, Kp, Kpc, dx_implicit, cpad_dx, Kp_dx} */
extern "C" int exvae_conv_plan(int N, int H, int W, int Cin, int KH, int KW, int stride, int pad, int ncat, int* plan) {
  EXVAE_CHECK_ARG(plan && N > 0 && H > 0 && W > 0 && Cin > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0 && ncat > 0);
  const ConvPlan c = conv_plan(N, H, W, Cin, KH, KW, stride, pad, ncat);
  EXVAE_CHECK_ARG(c.OH > 0 && c.OW > 0);
  plan[0] = c.OH; plan[1] = c.OW; plan[2] = c.implicit; plan[3] = c.cpad; plan[4] = c.Kp; plan[5] = c.Kpc;
  plan[6] = c.dx_implicit; plan[7] = c.cpad_dx; plan[8] = c.Kp_dx;
  return EXVAE_OK;
}

extern "C" size_t exvae_conv2d_fwd_workspace_bytes(int N, int H, int W, int Cin, int KH, int KW, int stride, int pad,
                                                   int ncat) {
  if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0) return 0;
  const ConvPlan c = conv_plan(N, H, W, Cin, KH, KW, stride, pad, ncat);
  return c.implicit ? 256 : align_up(sizeof(float) * (size_t)c.R * c.Kp, 256);
}

extern "C" int exvae_conv2d_fwd(const float* x, const float* wpk, const float* b0, const float* b1, int N, int H, int W,
                                int Cin, int KH, int KW, int stride, int pad, int O, int gated, int act, float lo,
                                float hi, float* out, float* sig, void* ws, size_t ws_bytes, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(x && wpk && out && N > 0 && H > 0 && W > 0 && Cin > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0 && O > 0);
  if (!tc_enabled()) return EXVAE_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  const int ncat = gated ? 2 * O : O;
  const ConvPlan c = conv_plan(N, H, W, Cin, KH, KW, stride, pad, ncat);
  EXVAE_CHECK_ARG(c.OH > 0 && c.OW > 0 && c.R < (1ll << 31));
  TcGemm g{};
  TcConv cv{};
  if (c.implicit) {
    cv.x = x; cv.N = N; cv.H = H; cv.W = W; cv.C = Cin; cv.KH = KH; cv.KW = KW; cv.stride = stride; cv.pad = pad;
    cv.OH = c.OH; cv.OW = c.OW; cv.bh = c.bh; cv.bn = c.bn;
    g.conv = &cv;
  } else {
    if (!ws || ws_bytes < sizeof(float) * (size_t)c.R * c.Kp) return EXVAE_ERR_WORKSPACE;
    float* col = static_cast<float*>(ws);
    int rc = conv_im2col(x, N, H, W, Cin, KH, KW, stride, pad, c.OH, c.OW, c.Kp, col, st);
    if (rc) return rc;
    g.a = col; g.a_rows = (int)c.R; g.a_cols = c.Kp;
  }
  g.a_mn = false;
  g.b = wpk; g.b_rows = ncat; g.b_cols = c.Kp; g.b_mn = false;
  g.M = (int)c.R; g.N = O; g.K = c.Kp; g.ldc = O; g.out0 = out;
  if (gated) {
    g.epi = TC_GATED; g.gated_O = O; g.bias0 = b0; g.bias1 = b1; g.out2 = sig;
  } else {
    g.epi = TC_BIAS_ACT; g.bias0 = b0; g.act = act; g.lo = lo; g.hi = hi;
  }
  return tc_gemm_launch(g, st);
}

extern "C" size_t exvae_conv2d_bwd_workspace_bytes(int N, int H, int W, int Cin, int KH, int KW, int stride, int pad, int O,
                                                   int gated) {
  if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || O <= 0) return 0;
  const int ncat = gated ? 2 * O : O;
  return conv_bwd_ws(conv_plan(N, H, W, Cin, KH, KW, stride, pad, ncat), ncat, O).bytes;
}

/* wbw: the backward operand the plan asks for: dx_implicit ? mode-1 packing [Cin][Kp_dx] : mode-0 packing with
 * cpad = Cin, [ncat][Kpc] (only read when dx != NULL). */
extern "C" int exvae_conv2d_bwd(const float* x, const float* wbw, const float* out, const float* sig, const float* dout,
                                int N, int H, int W, int Cin, int KH, int KW, int stride, int pad, int O, int gated,
                                int act, float lo, float hi, float* dx, float* dW0, float* db0, float* dW1, float* db1,
                                void* ws, size_t ws_bytes, int accumulate, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(x && dout && dW0 && ws && N > 0 && H > 0 && W > 0 && Cin > 0 && KH > 0 && KW > 0 && O > 0);
  EXVAE_CHECK_ARG(!gated || (out && sig && dW1));
  EXVAE_CHECK_ARG(!dx || wbw);
  if (!tc_enabled()) return EXVAE_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  const int ncat = gated ? 2 * O : O;
  const ConvPlan c = conv_plan(N, H, W, Cin, KH, KW, stride, pad, ncat);
  const ConvBwdWs wl = conv_bwd_ws(c, ncat, O);
  if (ws_bytes < wl.bytes) return EXVAE_ERR_WORKSPACE;
  char* w = static_cast<char*>(ws);
  float* dcat = reinterpret_cast<float*>(w + wl.off_dcat);
  float* cs = reinterpret_cast<float*>(w + wl.off_cs);
  float* col = reinterpret_cast<float*>(w + wl.off_col);
  float* part = reinterpret_cast<float*>(w + wl.off_part);
  const int R = (int)c.R;
  // 1. pre-activation gradient dcat [R][ldd] (+ per-block column sums for the bias gradients)
  if (c.ldd != ncat) EXVAE_CUDA(cudaMemsetAsync(dcat, 0, sizeof(float) * (size_t)R * c.ldd, st));
  {
    const TcBwdPlan& gp = wl.gp;
    const size_t sh = sizeof(float) * gp.RL * ncat;
    if (gated) {
      auto k = (O % 4 == 0) ? dpre_colsum_kernel<0, 4> : dpre_colsum_kernel<0, 1>;
      k<<<gp.S2, 256, sh, st>>>(dout, out, sig, R, O, 0, 0.f, 0.f, gp.RL, gp.iters, dcat, cs, c.ldd);
    } else {
      auto k = (O % 4 == 0) ? dpre_colsum_kernel<1, 4> : dpre_colsum_kernel<1, 1>;
      k<<<gp.S2, 256, sh, st>>>(dout, nullptr, out, R, O, act, lo, hi, gp.RL, gp.iters, dcat, cs, c.ldd);
    }
    EXVAE_CUDA(cudaGetLastError());
  }
  int rc;
  // 2. weight gradient.
  // Stride 1 and Cin % 32 == 0: IMPLICIT.  Both operands are copied into one zero-padded frame [N][Hp][Wp] (Hp = H + 2 pad,
  // output pixel (oy, ox) at frame position (oy, ox), input pixel (iy, ix) at (iy + pad, ix + pad)); then the input pixel of
  // tap (ky, kx) sits at the CONSTANT row offset ky*Wp + kx from the output pixel, every out-of-image read lands in the
  // zero border, and dW[co][tap][ci] = sum_q dYp[q][co] * Xp[q + ky*Wp + kx][ci] is a plain split-K GEMM whose B tile is
  // loaded with a per-tap row shift (TcGemm::dwc_*).  No patch matrix: 2 x 1.15 x (in + dcat) of copies instead of
  // 2 x taps x in.
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const long long Q = (long long)N * Hp * Wp;
  static const bool dw_implicit_on = [] { const char* e = getenv("EXVAE_CONV_DW_IMPLICIT"); return !(e && strcmp(e, "0") == 0); }();
  const bool dw_implicit = dw_implicit_on && stride == 1 && Cin % 32 == 0 && c.Kpc == c.taps * Cin && Q < (1ll << 31) &&
                           c.OH == Hp - KH + 1 && c.OW == Wp - KW + 1 &&
                           (size_t)Q * (c.ldd + Cin) <= (size_t)R * c.Kpc;      // the frames fit into the patch region
  if (dw_implicit) {
    float* dyp = col;                                   // [Q][ldd]
    float* xp = col + (size_t)Q * c.ldd;                // [Q][Cin]   (Q * ldd * 4 bytes is a multiple of 16)
    rc = conv_pad_frame(dcat, N, c.OH, c.OW, c.ldd, Hp, Wp, 0, 0, dyp, st);
    if (rc) return rc;
    rc = conv_pad_frame(x, N, H, W, Cin, Hp, Wp, pad, pad, xp, st);
    if (rc) return rc;
    const int kchunk = ceil_div(ceil_div((int)Q, wl.gp.S), 32) * 32;
    const int S = ceil_div((int)Q, kchunk);             // <= the planned split count: the partial planes fit
    TcGemm g{};
    g.a = dyp; g.a_rows = (int)Q; g.a_cols = c.ldd; g.a_mn = true;
    g.b = xp; g.b_rows = (int)Q; g.b_cols = Cin; g.b_mn = true;
    g.M = ncat; g.N = c.Kpc; g.K = (int)Q; g.epi = TC_SPLITK; g.out0 = part; g.ldc = c.Kpc;
    g.splits = S; g.kchunk = kchunk;
    g.dwc_cin = Cin; g.dwc_kw = KW; g.dwc_wp = Wp;
    rc = tc_gemm_launch(g, st);
    if (rc) return rc;
    rc = conv_unpack_wgrad(part, S, ncat, c.Kpc, Cin, O, Cin, KH, KW, dW0, dW1, cs, wl.gp.S2, db0, db1, accumulate, st);
    if (rc) return rc;
  } else {
    // patches recomputed (never saved by the forward), dWcat = dcat^T . col, split over the pixels
    rc = conv_im2col(x, N, H, W, Cin, KH, KW, stride, pad, c.OH, c.OW, c.Kpc, col, st);
    if (rc) return rc;
    TcGemm g{};
    g.a = dcat; g.a_rows = R; g.a_cols = c.ldd; g.a_mn = true;
    g.b = col; g.b_rows = R; g.b_cols = c.Kpc; g.b_mn = true;
    g.M = ncat; g.N = c.Kpc; g.K = R; g.epi = TC_SPLITK; g.out0 = part; g.ldc = c.Kpc;
    g.splits = wl.gp.S; g.kchunk = wl.gp.kchunk;
    rc = tc_gemm_launch(g, st);
    if (rc) return rc;
    rc = conv_unpack_wgrad(part, wl.gp.S, ncat, c.Kpc, Cin, O, Cin, KH, KW, dW0, dW1, cs, wl.gp.S2, db0, db1, accumulate, st);
    if (rc) return rc;
  }
  // 3. input gradient
  if (dx) {
    if (c.dx_implicit) {
      // stride 1: dx = conv(dcat, flipped filters) with padding K-1-pad, again an implicit GEMM
      TcConv cv{};
      cv.x = dcat; cv.N = N; cv.H = c.OH; cv.W = c.OW; cv.C = ncat; cv.KH = KH; cv.KW = KW; cv.stride = 1;
      cv.pad = KH - 1 - pad; cv.OH = H; cv.OW = W; cv.bh = c.bh_dx; cv.bn = c.bn_dx;
      TcGemm g{};
      g.conv = &cv; g.a_mn = false;
      g.b = wbw; g.b_rows = Cin; g.b_cols = c.Kp_dx; g.b_mn = false;
      g.M = N * H * W; g.N = Cin; g.K = c.Kp_dx; g.epi = TC_BIAS_ACT; g.act = EXVAE_ACT_NONE; g.out0 = dx; g.ldc = Cin;
      rc = tc_gemm_launch(g, st);
      if (rc) return rc;
    } else {
      // dcol [R][Kpc] = dcat . Wcat (A K-major, B MN-major), then the adjoint of im2col (gather form, no atomics)
      TcGemm g{};
      g.a = dcat; g.a_rows = R; g.a_cols = c.ldd; g.a_mn = false;
      g.b = wbw; g.b_rows = ncat; g.b_cols = c.Kpc; g.b_mn = true;
      g.M = R; g.N = c.Kpc; g.K = c.ldd; g.epi = TC_PLAIN; g.out0 = col; g.ldc = c.Kpc;
      rc = tc_gemm_launch(g, st);
      if (rc) return rc;
      rc = conv_col2im(col, N, H, W, Cin, KH, KW, stride, pad, c.OH, c.OW, c.Kpc, dx, st);
      if (rc) return rc;
    }
  }
  return EXVAE_OK;
}
