// Materialising distance primitives and the kNN exemplar selection (K2), sm_100a.
//
//   pairwise_distance                 utils/distributions.py:12-18   (fp64 inside, fp32 out)
//   log_normal_diag_vectorized        utils/distributions.py:21-25
//   log_p_z_exemplar, sum=False       models/BaseModel.py:98-109
//   (the fused distance + top-k selection of models/BaseModel.py:263-264 / utils/knn_on_latent.py:4-9 is knn_fused.cu)
//   torch.unique(nearest)             models/BaseModel.py:265
//
// The reference computes ||z||^2 + ||mu||^2 - 2 z.mu in fp64 and rounds to fp32 BEFORE the
// top-k, so the selected indices depend on that exact rounding.  The distance tile here is
// therefore accumulated with fp64 FMAs (products of fp32 inputs are exact in fp64; B200 keeps
// a full-rate fp64 pipe), combined in the reference's operation order and rounded once.
// Selection key is (distance_fp32, position) ascending: ties go to the lowest position.
#include "common.cuh"

namespace exvae {
namespace {

constexpr int PD_T = 64;   // tile edge
constexpr int PD_K = 16;   // k chunk
constexpr int PD_P = PD_T + 1;

enum PdMode { PD_PLAIN = 0, PD_LOGNORMAL = 1, PD_EXEMPLAR = 2, PD_EUCLID32 = 3 };

// 64x64 tile per CTA, 256 threads, 4x4 outputs per thread.
template <int MODE>
__global__ void __launch_bounds__(256) pairdist_kernel(const float* __restrict__ z, const float* __restrict__ mu,
                                                       const float* __restrict__ logvar,
                                                       const int64_t* __restrict__ z_idx,
                                                       const int64_t* __restrict__ mu_idx,
                                                       const int* __restrict__ row_counts, int B, int C, int D,
                                                       float* __restrict__ out0, float* __restrict__ out1) {
  __shared__ double As[PD_K][PD_P];
  __shared__ double Bs[PD_K][PD_P];
  __shared__ double na[PD_T], nb[PD_T];
  __shared__ float s_cst;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * PD_T, n0 = blockIdx.x * PD_T;
  constexpr bool SCALED = (MODE == PD_LOGNORMAL || MODE == PD_EXEMPLAR);

  if (SCALED && tid < 32) {
    float c = 0.f;
    for (int d = tid; d < D; d += 32) c += logvar[d] + kLog2Pi;
    c = warp_sum(c);
    if (tid == 0) s_cst = -0.5f * c;
  }
  double acc[4][4];
  float accf[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[i][j] = 0.0;
      accf[i][j] = 0.f;
    }
  double nrm = 0.0;  // tid < 64: row norms; 64 <= tid < 128: column norms

  for (int k0 = 0; k0 < D; k0 += PD_K) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + 256 * e;
      const int r = idx >> 4, kk = idx & 15;
      const int d = k0 + kk;
      float sg = 1.f;
      if (SCALED && d < D) sg = expf(0.5f * logvar[d]);
      float a = 0.f, b = 0.f;
      if (d < D && m0 + r < B) a = z[(size_t)(m0 + r) * D + d];
      if (d < D && n0 + r < C) b = mu[(size_t)(n0 + r) * D + d];
      if (SCALED) {
        a = a / sg;
        b = b / sg;
      }
      As[kk][r] = (double)a;
      Bs[kk][r] = (double)b;
    }
    __syncthreads();
    if (MODE != PD_EUCLID32) {
      if (tid < PD_T) {
#pragma unroll
        for (int kk = 0; kk < PD_K; ++kk) nrm = fma(As[kk][tid], As[kk][tid], nrm);
      } else if (tid < 2 * PD_T) {
#pragma unroll
        for (int kk = 0; kk < PD_K; ++kk) nrm = fma(Bs[kk][tid - PD_T], Bs[kk][tid - PD_T], nrm);
      }
    }
#pragma unroll
    for (int kk = 0; kk < PD_K; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx + 16 * j];
      if (MODE == PD_EUCLID32) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float df = (float)a[i] - (float)b[j];
            accf[i][j] += df * df;  // (z-mu)**2 summed in fp32, utils/knn_on_latent.py:7-8
          }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
  if (tid < PD_T) na[tid] = nrm;
  else if (tid < 2 * PD_T) nb[tid - PD_T] = nrm;
  __syncthreads();

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty + 16 * i, b = m0 + r;
    if (b >= B) continue;
    float logden = 0.f;
    long long zi = 0;
    if (MODE == PD_EXEMPLAR) {
      logden = logf((float)C - (float)(row_counts ? row_counts[b] : 0));
      zi = z_idx ? z_idx[b] : 0;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cc = tx + 16 * j, c = n0 + cc;
      if (c >= C) continue;
      const size_t o = (size_t)b * C + c;
      if (MODE == PD_EUCLID32) {
        out0[o] = sqrtf(accf[i][j]);
        continue;
      }
      const float pd = (float)((na[r] + nb[cc]) + (-2.0 * acc[i][j]));
      if (MODE == PD_PLAIN) {
        out0[o] = pd;
      } else if (MODE == PD_LOGNORMAL) {
        out0[o] = s_cst - 0.5f * pd;
        if (out1) out1[o] = pd;
      } else {
        float p = s_cst - 0.5f * pd;
        if (z_idx && mu_idx && mu_idx[c] == zi) p = -INFINITY;
        out0[o] = p - logden;
      }
    }
  }
}

// row_counts[b] = #{n : mu_idx[n] == z_idx[b]}   (models/BaseModel.py:104-107)
__global__ void __launch_bounds__(256) mask_count_kernel(const int64_t* __restrict__ z_idx,
                                                         const int64_t* __restrict__ mu_idx, int B, int C,
                                                         int* __restrict__ row_counts) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const long long zi = z_idx[b];
  int c = 0;
  for (int n = lane; n < C; n += 32) c += (mu_idx[n] == zi);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) row_counts[b] = c;
}

// ---------------------------------------------------------------------------------- top-k
struct Cand {
  float v;
  int i;
};
__device__ __forceinline__ bool cand_less(float v, int i, float bv, int bi) { return v < bv || (v == bv && i < bi); }
__device__ __forceinline__ Cand cand_min(Cand a, Cand b) { return cand_less(b.v, b.i, a.v, a.i) ? b : a; }

// Merge G per-shard lists: one warp per row over G*k candidates keyed by (dist, global position).
__global__ void __launch_bounds__(256) knn_merge_kernel(const int64_t* __restrict__ idx, const float* __restrict__ dist,
                                                        int G, int B, int k, int64_t* __restrict__ out_idx,
                                                        float* __restrict__ out_dist) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int n = G * k;
  float lv = -INFINITY;
  long long li = -1;
  for (int p = 0; p < k; ++p) {
    float bv = INFINITY;
    long long bi = INT64_MAX;
    for (int c = lane; c < n; c += 32) {
      const int g = c / k, j = c - g * k;
      const size_t o = ((size_t)g * B + b) * k + j;
      const float v = dist[o];
      const long long i = idx[o];
      if (i < 0) continue;
      const bool after = (v > lv) || (v == lv && i > li);
      if (after && (v < bv || (v == bv && i < bi))) {
        bv = v;
        bi = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov < bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      out_idx[(size_t)b * k + p] = bi == INT64_MAX ? -1 : bi;
      out_dist[(size_t)b * k + p] = bv;
    }
    lv = bv;
    li = bi;
  }
}

// Sorted unique of positions in [0, range): flag scatter + single-CTA scan/compaction.
__global__ void __launch_bounds__(1024) unique_positions_kernel(const int64_t* __restrict__ pos, int n, int range,
                                                                int64_t* __restrict__ out, int* __restrict__ out_count,
                                                                int* __restrict__ flags) {
  __shared__ int wsum[32];
  __shared__ int s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < range; i += 1024) flags[i] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += 1024) {
    const long long p = pos[i];
    if (p >= 0 && p < range) flags[p] = 1;
  }
  __syncthreads();
  const int per = ceil_div(range, 1024);
  const int lo = tid * per, hi = min(range, lo + per);
  int cnt = 0;
  for (int i = lo; i < hi; ++i) cnt += flags[i];
  int inc = cnt;  // inclusive warp scan
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int v = wsum[lane];
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    wsum[lane] = s - v;  // exclusive
    if (lane == 31) s_total = s;
  }
  __syncthreads();
  int w = wsum[warp] + inc - cnt;
  for (int i = lo; i < hi; ++i)
    if (flags[i]) out[w++] = i;
  if (tid == 0) *out_count = s_total;
  // fixed-size result (graph-capturable kNN mode): the tail [count, n) repeats the first selected position, so
  // every entry is a valid row to gather / encode; consumers ignore it through the device-side count
  __syncthreads();
  const long long first = s_total > 0 ? out[0] : 0;
  for (int i = s_total + tid; i < n; i += 1024) out[i] = first;
}

// ------------------------------------------------------------------------------ row movement
// Grid = 4 CTAs per SM, all resident at once: a launch with PENDING CTAs holds back every later kernel of the step (the
// block scheduler dispatches in order), and this gather runs next to the optimizer kernels.  Four independent 16-byte
// loads per thread and iteration keep ~10 MB in flight, enough for the HBM latency-bandwidth product.
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx,
                                                          long long n_rows, int row_len, int vec,
                                                          float* __restrict__ out) {
  const long long per_row = row_len / vec;
  const long long total = n_rows * per_row;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (vec == 4) {
    const float4* src4 = reinterpret_cast<const float4*>(src);
    float4* out4 = reinterpret_cast<float4*>(out);
    for (; e + 3 * stride < total; e += 4 * stride) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long ee = e + u * stride;
        const long long r = ee / per_row, c = ee - r * per_row;
        v[u] = src4[idx[r] * per_row + c];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) out4[e + u * stride] = v[u];
    }
    for (; e < total; e += stride) {
      const long long r = e / per_row, c = e - r * per_row;
      out4[e] = src4[idx[r] * per_row + c];
    }
  } else {
    for (; e < total; e += stride) {
      const long long r = e / per_row, c = e - r * per_row;
      out[e] = src[idx[r] * row_len + c];
    }
  }
}
__global__ void __launch_bounds__(256) scatter_rows_kernel(float* __restrict__ dst, const int64_t* __restrict__ idx,
                                                           long long n_rows, int row_len, int vec,
                                                           const float* __restrict__ src) {
  const long long per_row = row_len / vec;
  const long long total = n_rows * per_row;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / per_row, v = e - r * per_row;
    const long long s = idx[r];
    if (vec == 4)
      reinterpret_cast<float4*>(dst + s * row_len)[v] = reinterpret_cast<const float4*>(src)[e];
    else
      dst[s * row_len + v] = src[e];
  }
}

template <int MODE>
int launch_pairdist(const float* z, const float* mu, const float* logvar, const int64_t* z_idx, const int64_t* mu_idx,
                    const int* row_counts, int B, int C, int D, float* out0, float* out1, cudaStream_t st) {
  dim3 grid(ceil_div(C, PD_T), ceil_div(B, PD_T));
  pairdist_kernel<MODE><<<grid, 256, 0, st>>>(z, mu, logvar, z_idx, mu_idx, row_counts, B, C, D, out0, out1);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? EXVAE_OK : (int)e;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace exvae

using namespace exvae;

extern "C" int exvae_pairwise_distance(const float* z, const float* means, int B, int C, int D, float* out,
                                       exvae_stream_t stream) {
  EXVAE_CHECK_ARG(z && means && out && B > 0 && C > 0 && D > 0);
  return launch_pairdist<PD_PLAIN>(z, means, nullptr, nullptr, nullptr, nullptr, B, C, D, out, nullptr,
                                   as_stream(stream));
}

extern "C" int exvae_log_normal_diag_vectorized(const float* x, const float* mean, const float* log_var, int B, int C,
                                                int D, float* log_normal, float* pair_dist, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(x && mean && log_var && log_normal && B > 0 && C > 0 && D > 0);
  return launch_pairdist<PD_LOGNORMAL>(x, mean, log_var, nullptr, nullptr, nullptr, B, C, D, log_normal, pair_dist,
                                       as_stream(stream));
}

extern "C" int exvae_prior_logprob_matrix(const float* z, const float* mu, const float* logvar, const int64_t* z_idx,
                                          const int64_t* mu_idx, int B, int C, int D, float* out, int* row_counts,
                                          exvae_stream_t stream) {
  EXVAE_CHECK_ARG(z && mu && logvar && out && B > 0 && C > 0 && D > 0);
  cudaStream_t st = as_stream(stream);
  const bool mask = z_idx && mu_idx;
  if (mask) {
    EXVAE_CHECK_ARG(row_counts != nullptr);
    mask_count_kernel<<<ceil_div(B, 8), 256, 0, st>>>(z_idx, mu_idx, B, C, row_counts);
    EXVAE_CUDA(cudaGetLastError());
  }
  return launch_pairdist<PD_EXEMPLAR>(z, mu, logvar, mask ? z_idx : nullptr, mask ? mu_idx : nullptr,
                                      mask ? row_counts : nullptr, B, C, D, out, nullptr, st);
}

extern "C" int exvae_knn_merge(const int64_t* idx, const float* dist, int G, int B, int k, int64_t* out_idx,
                               float* out_dist, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(idx && dist && out_idx && out_dist && G > 0 && B > 0 && k > 0);
  knn_merge_kernel<<<ceil_div(B, 8), 256, 0, as_stream(stream)>>>(idx, dist, G, B, k, out_idx, out_dist);
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_unique_positions(const int64_t* pos, int n, int range, int64_t* out, int* out_count, int* flags,
                                      exvae_stream_t stream) {
  EXVAE_CHECK_ARG(pos && out && out_count && flags && n > 0 && range > 0);
  unique_positions_kernel<<<1, 1024, 0, as_stream(stream)>>>(pos, n, range, out, out_count, flags);
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_gather_rows(const float* src, const int64_t* idx, int n_rows, int row_len, float* out,
                                 exvae_stream_t stream) {
  EXVAE_CHECK_ARG(src && idx && out && n_rows > 0 && row_len > 0);
  const int vec = (row_len % 4 == 0 && aligned16(src) && aligned16(out)) ? 4 : 1;
  const long long total = (long long)n_rows * (row_len / vec);
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 4);
  gather_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(src, idx, n_rows, row_len, vec, out);
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_scatter_rows(float* dst, const int64_t* idx, int n_rows, int row_len, const float* src,
                                  exvae_stream_t stream) {
  EXVAE_CHECK_ARG(dst && idx && src && n_rows > 0 && row_len > 0);
  const int vec = (row_len % 4 == 0 && aligned16(src) && aligned16(dst)) ? 4 : 1;
  const long long total = (long long)n_rows * (row_len / vec);
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
  scatter_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(dst, idx, n_rows, row_len, vec, src);
  EXVAE_RETURN_LAST_ERROR();
}
