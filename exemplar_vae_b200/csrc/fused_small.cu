// Small fused kernels that keep every launch of the captured training step inside this library (no ATen fill /
// add / cat / copy kernels on the step):
//   exvae_zero                 cudaMemsetAsync of the flat gradient buffer
//   exvae_bcast_scalar_*       prior_log_variance [1] -> [D] row (models/BaseModel.py:212-214) and its gradient sum
//   exvae_reparam_logq_*       z = mu + exp(lv/2) eps  and  log q(z|x) in ONE pass (models/BaseModel.py:79-82 +
//                              utils/distributions.py:28-33), same operation order as the two separate kernels
//   exvae_concat_cols_*        torch.cat((a, b), 1) of models/AbsHModel.py:55,83 and its split backward
#include "common.cuh"

namespace exvae {
namespace {

inline int ew_blocks(long long n) { return (int)std::min<long long>((n + 255) / 256, 148LL * 16); }

__global__ void __launch_bounds__(128) bcast_scalar_kernel(const float* __restrict__ src, int n, float* __restrict__ out) {
  const float v = src[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = v;
}
// out[0] (+)= sum_i src[i], one block, fixed order per thread + shuffle tree (deterministic)
__global__ void __launch_bounds__(128) sum_to_scalar_kernel(const float* __restrict__ src, int n, float* __restrict__ out,
                                                            int accumulate) {
  __shared__ float sh[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float a = 0.f;
  for (int i = tid; i < n; i += 128) a += src[i];
  a = warp_sum(a);
  if (lane == 0) sh[warp] = a;
  __syncthreads();
  if (tid == 0) {
    const float t = (sh[0] + sh[1]) + (sh[2] + sh[3]);
    out[0] = accumulate ? out[0] + t : t;
  }
}

// one warp per row
__global__ void __launch_bounds__(256) reparam_logq_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                                               const float* __restrict__ eps, int B, int D,
                                                               float* __restrict__ z, float* __restrict__ logq) {
  const int lane = threadIdx.x & 31, b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  float a = 0.f;
  for (int d = lane; d < D; d += 32) {
    const size_t o = (size_t)b * D + d;
    const float l = lv[o], m = mu[o];
    const float zv = __fadd_rn(__fmul_rn(eps[o], expf(0.5f * l)), m);      // reparam_fwd_kernel
    z[o] = zv;
    const float df = zv - m;                                                // lognormal_fwd_kernel on the rounded z
    a += -0.5f * (l + kLog2Pi + df * df / expf(l));
  }
  a = warp_sum(a);
  if (lane == 0) logq[b] = a;
}
// dz: gradient reaching z from its OTHER consumers (decoder, prior; may be NULL), dlogq [B] (may be NULL).
// Same arithmetic as lognormal_bwd_kernel + the dz accumulation + reparam_bwd_kernel.
__global__ void __launch_bounds__(256) reparam_logq_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                                               const float* __restrict__ eps,
                                                               const float* __restrict__ z,
                                                               const float* __restrict__ dz,
                                                               const float* __restrict__ dlogq, long long n, int D,
                                                               float* __restrict__ dmu, float* __restrict__ dlv) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float g = dlogq ? dlogq[e / D] : 0.f;
    const float l = lv[e];
    const float df = z[e] - mu[e], iv = 1.f / expf(l);
    const float t = g * df * iv;
    const float dzt = (dz ? dz[e] : 0.f) - t;          // total gradient of z
    if (dmu) dmu[e] = dzt + t;
    if (dlv) dlv[e] = dzt * eps[e] * expf(0.5f * l) * 0.5f + g * (-0.5f + 0.5f * df * df * iv);
  }
}

// one warp per row
__global__ void __launch_bounds__(256) weight_norm_fwd_kernel(const float* __restrict__ v, const float* __restrict__ g, int R,
                                                              int K, float* __restrict__ w) {
  const int lane = threadIdx.x & 31, r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  float ss = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float x = v[(size_t)r * K + k];
    ss = fmaf(x, x, ss);
  }
  ss = warp_sum(ss);
  const float sc = g[r] / sqrtf(ss);
  for (int k = lane; k < K; k += 32) w[(size_t)r * K + k] = v[(size_t)r * K + k] * sc;
}
// w = g v / n,  n = ||v||:   dg = <dw, v> / n ;  dv = (g / n) (dw - v <dw, v> / n^2)
__global__ void __launch_bounds__(256) weight_norm_bwd_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                              const float* __restrict__ dw, int R, int K,
                                                              float* __restrict__ dv, float* __restrict__ dg,
                                                              int accumulate) {
  const int lane = threadIdx.x & 31, r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  float ss = 0.f, dot = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float x = v[(size_t)r * K + k];
    ss = fmaf(x, x, ss);
    dot = fmaf(dw[(size_t)r * K + k], x, dot);
  }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  const float n = sqrtf(ss), gn = g[r] / n, c = dot / ss;
  for (int k = lane; k < K; k += 32) {
    const size_t o = (size_t)r * K + k;
    const float t = gn * (dw[o] - v[o] * c);
    dv[o] = accumulate ? dv[o] + t : t;
  }
  if (lane == 0) dg[r] = accumulate ? dg[r] + dot / n : dot / n;
}

__global__ void __launch_bounds__(256) gather_index_kernel(const long long* __restrict__ src, const long long* __restrict__ idx,
                                                           int n, long long* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = src[idx[i]];
}

__global__ void __launch_bounds__(256) concat_cols_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                              long long R, int Ka, int Kb, float* __restrict__ out) {
  const int K = Ka + Kb;
  const long long n = R * K;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / K;
    const int c = (int)(e - r * K);
    out[e] = c < Ka ? a[r * Ka + c] : b[r * Kb + (c - Ka)];
  }
}
__global__ void __launch_bounds__(256) concat_cols_bwd_kernel(const float* __restrict__ dout, long long R, int Ka, int Kb,
                                                              float* __restrict__ da, float* __restrict__ db) {
  const int K = Ka + Kb;
  const long long n = R * K;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / K;
    const int c = (int)(e - r * K);
    if (c < Ka) {
      if (da) da[r * Ka + c] = dout[e];
    } else if (db) {
      db[r * Kb + (c - Ka)] = dout[e];
    }
  }
}

}  // namespace
}  // namespace exvae

using namespace exvae;

extern "C" int exvae_zero(void* ptr, size_t bytes, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(ptr != nullptr);
  if (bytes == 0) return EXVAE_OK;
  EXVAE_CUDA(cudaMemsetAsync(ptr, 0, bytes, as_stream(stream)));
  return EXVAE_OK;
}

extern "C" int exvae_bcast_scalar(const float* src, int n, float* out, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(src && out && n > 0);
  bcast_scalar_kernel<<<ceil_div(n, 128) > 64 ? 64 : ceil_div(n, 128), 128, 0, as_stream(stream)>>>(src, n, out);
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_sum_to_scalar(const float* src, int n, float* out, int accumulate, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(src && out && n > 0);
  sum_to_scalar_kernel<<<1, 128, 0, as_stream(stream)>>>(src, n, out, accumulate);
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_reparam_logq_fwd(const float* mu, const float* logvar, const float* eps, int B, int D, float* z,
                                      float* logq, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(mu && logvar && eps && z && logq && B > 0 && D > 0);
  reparam_logq_fwd_kernel<<<ceil_div(B, 8), 256, 0, as_stream(stream)>>>(mu, logvar, eps, B, D, z, logq);
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_reparam_logq_bwd(const float* mu, const float* logvar, const float* eps, const float* z,
                                      const float* dz, const float* dlogq, int B, int D, float* dmu, float* dlogvar,
                                      exvae_stream_t stream) {
  EXVAE_CHECK_ARG(mu && logvar && eps && z && B > 0 && D > 0 && (dmu || dlogvar));
  const long long n = (long long)B * D;
  reparam_logq_bwd_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(mu, logvar, eps, z, dz, dlogq, n, D, dmu, dlogvar);
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_weight_norm_fwd(const float* v, const float* g, int R, int K, float* w, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(v && g && w && R > 0 && K > 0);
  weight_norm_fwd_kernel<<<ceil_div(R, 8), 256, 0, as_stream(stream)>>>(v, g, R, K, w);
  EXVAE_RETURN_LAST_ERROR();
}
extern "C" int exvae_weight_norm_bwd(const float* v, const float* g, const float* dw, int R, int K, float* dv, float* dg,
                                     int accumulate, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(v && g && dw && dv && dg && R > 0 && K > 0);
  weight_norm_bwd_kernel<<<ceil_div(R, 8), 256, 0, as_stream(stream)>>>(v, g, dw, R, K, dv, dg, accumulate);
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_gather_index(const int64_t* src, const int64_t* idx, int n, int64_t* out, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(src && idx && out && n > 0);
  gather_index_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(reinterpret_cast<const long long*>(src),
                                                                   reinterpret_cast<const long long*>(idx), n,
                                                                   reinterpret_cast<long long*>(out));
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_concat_cols_fwd(const float* a, const float* b, int64_t R, int Ka, int Kb, float* out,
                                     exvae_stream_t stream) {
  EXVAE_CHECK_ARG(a && b && out && R > 0 && Ka > 0 && Kb > 0);
  concat_cols_fwd_kernel<<<ew_blocks(R * (Ka + Kb)), 256, 0, as_stream(stream)>>>(a, b, R, Ka, Kb, out);
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_concat_cols_bwd(const float* dout, int64_t R, int Ka, int Kb, float* da, float* db,
                                     exvae_stream_t stream) {
  EXVAE_CHECK_ARG(dout && (da || db) && R > 0 && Ka > 0 && Kb > 0);
  concat_cols_bwd_kernel<<<ew_blocks(R * (Ka + Kb)), 256, 0, as_stream(stream)>>>(dout, R, Ka, Kb, da, db);
  EXVAE_RETURN_LAST_ERROR();
}
