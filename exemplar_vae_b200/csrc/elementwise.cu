// Element-wise / row-reduction pieces of the ELBO, the counter-based RNG and the fused
// AdamNormGrad step (sm_100a).  All are HBM-bound streaming kernels: coalesced accesses,
// one warp per row for the [B] reductions, no shared-memory staging needed.
//
//   reparameterize             models/BaseModel.py:79-82
//   log_normal_diag            utils/distributions.py:28-33
//   log_normal_standard        utils/distributions.py:36-41
//   log_bernoulli              utils/distributions.py:44-51
//   log_logistic_256           utils/distributions.py:54-66
//   loss = -RE + beta*KL, mean models/BaseModel.py:71-75
//   AdamNormGrad.step          utils/optimizer.py:32-80
//   bernoulli / randint / normal draws   utils/training.py:31, models/BaseModel.py:245,257,81
#include <algorithm>

#include "common.cuh"

namespace exvae {
namespace {

constexpr float kMinEps = 1e-5f;
constexpr float kMaxEps = 1.f - 1e-5f;

inline int ew_blocks(long long n) { return (int)std::min<long long>((n + 255) / 256, 148LL * 16); }

__global__ void __launch_bounds__(256) reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                                          const float* __restrict__ eps, long long n,
                                                          float* __restrict__ z) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    z[e] = __fadd_rn(__fmul_rn(eps[e], expf(0.5f * lv[e])), mu[e]);
}
__global__ void __launch_bounds__(256) reparam_bwd_kernel(const float* __restrict__ lv, const float* __restrict__ eps,
                                                          const float* __restrict__ dz, long long n,
                                                          float* __restrict__ dmu, float* __restrict__ dlv) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float g = dz[e];
    if (dmu) dmu[e] = g;
    if (dlv) dlv[e] = g * eps[e] * expf(0.5f * lv[e]) * 0.5f;
  }
}

// ---- one warp per row ---------------------------------------------------------------
__global__ void __launch_bounds__(256) lognormal_fwd_kernel(const float* __restrict__ x, const float* __restrict__ m,
                                                            const float* __restrict__ lv, int B, int D,
                                                            float* __restrict__ out) {
  const int lane = threadIdx.x & 31, b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  float a = 0.f;
  for (int d = lane; d < D; d += 32) {
    const size_t o = (size_t)b * D + d;
    const float df = x[o] - m[o], l = lv[o];
    a += -0.5f * (l + kLog2Pi + df * df / expf(l));
  }
  a = warp_sum(a);
  if (lane == 0) out[b] = a;
}
__global__ void __launch_bounds__(256) lognormal_bwd_kernel(const float* __restrict__ x, const float* __restrict__ m,
                                                            const float* __restrict__ lv,
                                                            const float* __restrict__ dout, long long n, int D,
                                                            float* __restrict__ dx, float* __restrict__ dm,
                                                            float* __restrict__ dlv) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float g = dout[e / D];
    const float df = x[e] - m[e], iv = 1.f / expf(lv[e]);
    const float t = g * df * iv;
    if (dx) dx[e] = -t;
    if (dm) dm[e] = t;
    if (dlv) dlv[e] = g * (-0.5f + 0.5f * df * df * iv);
  }
}
__global__ void __launch_bounds__(256) lognormstd_fwd_kernel(const float* __restrict__ x, int B, int D,
                                                             float* __restrict__ out) {
  const int lane = threadIdx.x & 31, b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  float a = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = x[(size_t)b * D + d];
    a += -0.5f * v * v - 0.5f * kLog2Pi;
  }
  a = warp_sum(a);
  if (lane == 0) out[b] = a;
}
__global__ void __launch_bounds__(256) lognormstd_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dout,
                                                             long long n, int D, float* __restrict__ dx) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    dx[e] = -dout[e / D] * x[e];
}

// one block of 128 threads per row: 512 rows x 784 pixels with two logf each is latency-bound with a warp per row
__global__ void __launch_bounds__(128) bernoulli_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                            int B, int P, float* __restrict__ out) {
  __shared__ float sh[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
  float a = 0.f;
  for (int d = tid; d < P; d += 128) {
    const size_t o = (size_t)b * P + d;
    const float p = fminf(fmaxf(mean[o], kMinEps), kMaxEps), xv = x[o];
    a += xv * logf(p) + (1.f - xv) * logf(1.f - p);
  }
  a = warp_sum(a);
  if (lane == 0) sh[warp] = a;
  __syncthreads();
  if (tid == 0) out[b] = (sh[0] + sh[1]) + (sh[2] + sh[3]);
}
__global__ void __launch_bounds__(256) bernoulli_bwd_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                            const float* __restrict__ dout, long long n, int P,
                                                            float* __restrict__ dmean) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float mv = mean[e], xv = x[e];
    float g = 0.f;
    if (mv >= kMinEps && mv <= kMaxEps) g = dout[e / P] * (xv / mv - (1.f - xv) / (1.f - mv));
    dmean[e] = g;
  }
}

__device__ __forceinline__ float sigm(float v) { return 1.f / (1.f + expf(-v)); }
__global__ void __launch_bounds__(256) logistic_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                           const float* __restrict__ lv, int B, int P,
                                                           float* __restrict__ out) {
  const int lane = threadIdx.x & 31, b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float bin = 1.f / 256.f;
  float a = 0.f;
  for (int d = lane; d < P; d += 32) {
    const size_t o = (size_t)b * P + d;
    const float sc = expf(lv[o]);
    const float xs = (floorf(x[o] / bin) * bin - mean[o]) / sc;
    a += logf(sigm(xs + bin / sc) - sigm(xs) + 1e-7f);
  }
  a = warp_sum(a);
  if (lane == 0) out[b] = a;
}
__global__ void __launch_bounds__(256) logistic_bwd_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                           const float* __restrict__ lv,
                                                           const float* __restrict__ dout, long long n, int P,
                                                           float* __restrict__ dmean, float* __restrict__ dlv) {
  const float bin = 1.f / 256.f;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float g = dout[e / P];
    const float sc = expf(lv[e]);
    const float xs = (floorf(x[e] / bin) * bin - mean[e]) / sc;
    const float bs = bin / sc;
    const float cp = sigm(xs + bs), cm = sigm(xs);
    const float it = 1.f / (cp - cm + 1e-7f);
    const float dp = cp * (1.f - cp), dm = cm * (1.f - cm);
    if (dmean) dmean[e] = g * (dp - dm) * it * (-1.f / sc);
    if (dlv) dlv[e] = g * (dp * (-xs - bs) + dm * xs) * it;
  }
}

// out3 = {mean(-RE + beta*KL), mean(RE), mean(KL)} or per-sample loss
__global__ void __launch_bounds__(1024) elbo_reduce_kernel(const float* __restrict__ RE, const float* __restrict__ KL,
                                                           int B, float beta, const float* __restrict__ beta_dev,
                                                           int average, float* __restrict__ out3,
                                                           float* __restrict__ loss_b) {
  __shared__ float sh[3][32];
  if (beta_dev) beta = *beta_dev;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float a = 0.f, r = 0.f, k = 0.f;
  for (int b = tid; b < B; b += blockDim.x) {
    const float re = RE[b], kl = KL[b];
    const float l = -re + beta * kl;
    if (loss_b) loss_b[b] = l;
    a += l;
    r += re;
    k += kl;
  }
  if (!average) return;
  a = warp_sum(a); r = warp_sum(r); k = warp_sum(k);
  if (lane == 0) { sh[0][warp] = a; sh[1][warp] = r; sh[2][warp] = k; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    a = lane < nw ? sh[0][lane] : 0.f;
    r = lane < nw ? sh[1][lane] : 0.f;
    k = lane < nw ? sh[2][lane] : 0.f;
    a = warp_sum(a); r = warp_sum(r); k = warp_sum(k);
    if (lane == 0) {
      const float inv = 1.f / (float)B;
      out3[0] = a * inv; out3[1] = r * inv; out3[2] = k * inv;
    }
  }
}

__global__ void __launch_bounds__(256) lincomb4_kernel(const float* __restrict__ x0, const float* __restrict__ x1,
                                                       const float* __restrict__ x2, const float* __restrict__ x3,
                                                       float c0, float c1, float c2, float c3, long long n,
                                                       float* __restrict__ out) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    float a = 0.f;
    if (x0) a += c0 * x0[e];
    if (x1) a += c1 * x1[e];
    if (x2) a += c2 * x2[e];
    if (x3) a += c3 * x3[e];
    out[e] = a;
  }
}

__global__ void __launch_bounds__(256) elbo_reduce_bwd_kernel(const float* __restrict__ g3,
                                                              const float* __restrict__ g_loss_b, int B, float beta,
                                                              const float* __restrict__ beta_dev, int average,
                                                              float* __restrict__ dRE, float* __restrict__ dKL) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (beta_dev) beta = *beta_dev;
  if (average) {
    const float inv = 1.f / (float)B;
    dRE[b] = (-g3[0] + g3[1]) * inv;
    dKL[b] = (beta * g3[0] + g3[2]) * inv;
  } else {
    const float g = g_loss_b[b];
    dRE[b] = -g;
    dKL[b] = beta * g;
  }
}

// ---- Philox4x32-10 ------------------------------------------------------------------
struct Philox {
  uint32_t c[4], k[2];
  __device__ Philox(uint64_t seed, uint64_t subseq, uint64_t offset) {
    k[0] = (uint32_t)seed; k[1] = (uint32_t)(seed >> 32);
    c[0] = (uint32_t)offset; c[1] = (uint32_t)(offset >> 32);
    c[2] = (uint32_t)subseq; c[3] = (uint32_t)(subseq >> 32);
  }
  __device__ uint4 next() {
    uint32_t c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], k0 = k[0], k1 = k[1];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    if (++c[0] == 0) ++c[1];
    return make_uint4(c0, c1, c2, c3);
  }
};
// Optional in-kernel advance of the draw counter: every block reads counter[0] before it does anything else and takes
// a ticket (counter[1], wraps back to 0) when it is done; the block that takes the last ticket therefore runs after
// all reads and bumps the counter, which saves the separate one-thread "advance" launch after each draw.
__device__ __forceinline__ void rng_finish(uint64_t* counter, int advance) {
  if (!advance || !counter) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicInc(reinterpret_cast<unsigned int*>(counter + 1), gridDim.x - 1);
    if (t == gridDim.x - 1) counter[0] += 1;
  }
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }  // [0,1)

__global__ void __launch_bounds__(256) rng_bernoulli_kernel(const float* __restrict__ p, long long n, uint64_t seed,
                                                            uint64_t* counter, uint64_t subseq, int advance,
                                                            float* __restrict__ out) {
  const uint64_t off = counter ? *counter : 0;
  const long long nq = (n + 3) / 4;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nq; q += (long long)gridDim.x * blockDim.x) {
    Philox ph(seed, (subseq << 40) + (uint64_t)q, off);
    const uint4 r = ph.next();
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long e = q * 4 + j;
      if (e < n) out[e] = u01(rr[j]) < p[e] ? 1.f : 0.f;
    }
  }
  rng_finish(counter, advance);
}
__global__ void __launch_bounds__(256) rng_normal_kernel(long long n, uint64_t seed, uint64_t* counter, uint64_t subseq,
                                                         int advance, float* __restrict__ out) {
  const uint64_t off = counter ? *counter : 0;
  const long long nq = (n + 3) / 4;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nq; q += (long long)gridDim.x * blockDim.x) {
    Philox ph(seed, (subseq << 40) + (uint64_t)q, off);
    const uint4 r = ph.next();
    float v[4];
    {
      const float u1 = 1.f - u01(r.x), u2 = u01(r.y);  // u1 in (0,1]
      const float rad = sqrtf(-2.f * logf(u1));
      float s, c;
      sincospif(2.f * u2, &s, &c);
      v[0] = rad * c; v[1] = rad * s;
    }
    {
      const float u1 = 1.f - u01(r.z), u2 = u01(r.w);
      const float rad = sqrtf(-2.f * logf(u1));
      float s, c;
      sincospif(2.f * u2, &s, &c);
      v[2] = rad * c; v[3] = rad * s;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long e = q * 4 + j;
      if (e < n) out[e] = v[j];
    }
  }
  rng_finish(counter, advance);
}
__global__ void __launch_bounds__(256) rng_randint_kernel(long long low, unsigned long long range, long long n,
                                                          uint64_t seed, uint64_t* counter, uint64_t subseq,
                                                          int advance, int64_t* __restrict__ out) {
  const uint64_t off = counter ? *counter : 0;
  const long long nq = (n + 1) / 2;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nq; q += (long long)gridDim.x * blockDim.x) {
    Philox ph(seed, (subseq << 40) + (uint64_t)q, off);
    const uint4 r = ph.next();
    const unsigned long long a = ((unsigned long long)r.x << 32) | r.y, b = ((unsigned long long)r.z << 32) | r.w;
    if (q * 2 < n) out[q * 2] = low + (long long)(a % range);
    if (q * 2 + 1 < n) out[q * 2 + 1] = low + (long long)(b % range);
  }
  rng_finish(counter, advance);
}
__global__ void rng_advance_kernel(uint64_t* counter, uint64_t by) { *counter += by; }

// ---- AdamNormGrad ------------------------------------------------------------------
struct AdamRec {
  float* p;
  float* g;
  float* m;
  float* v;
  long long n;
};
// ||g||_2^2 partials: grid = (tensors, ADAM_NCH chunks), deterministic tree per chunk; block (0,0) also
// bumps the step counter.  The update kernel adds the ADAM_NCH partials in a fixed order.
constexpr int ADAM_NCH = 16;
__global__ void __launch_bounds__(256) adam_norm_kernel(const AdamRec* __restrict__ table, float* __restrict__ norms,
                                                        int64_t* __restrict__ step, float lr, float beta1, float beta2,
                                                        float* __restrict__ step_size_out) {
  __shared__ float sh[8];
  const AdamRec rec = table[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float a = 0.f;
  if (rec.g)
    for (long long e = (long long)blockIdx.y * blockDim.x + tid; e < rec.n; e += (long long)ADAM_NCH * blockDim.x) {
      const float g = rec.g[e];
      a = fmaf(g, g, a);
    }
  a = warp_sum(a);
  if (lane == 0) sh[warp] = a;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i];
    norms[blockIdx.x * ADAM_NCH + blockIdx.y] = t;
    if (blockIdx.x == 0 && blockIdx.y == 0) {
      // bump the step counter and evaluate the bias-corrected step size ONCE (fp64 pow), not in every update thread
      const long long ts = step[0] + 1;
      step[0] = ts;
      const double bc1 = 1.0 - pow((double)beta1, (double)ts), bc2 = 1.0 - pow((double)beta2, (double)ts);
      *step_size_out = (float)((double)lr * sqrt(bc2) / bc1);
    }
  }
}
__global__ void __launch_bounds__(256) adam_update_kernel(const AdamRec* __restrict__ table,
                                                          const float* __restrict__ norms,
                                                          const float* __restrict__ step_size_in, float beta1,
                                                          float beta2, float eps, float wd) {
  const AdamRec rec = table[blockIdx.y];
  if (!rec.g) return;
  const long long base = (long long)blockIdx.x * blockDim.x * 4;
  if (base >= rec.n) return;
  const float step_size = *step_size_in;
  float nsq = 0.f;
#pragma unroll
  for (int i = 0; i < ADAM_NCH; ++i) nsq += norms[blockIdx.y * ADAM_NCH + i];
  const float inv = 1.f / (sqrtf(nsq) + 1e-7f);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long e = base + (long long)j * blockDim.x + threadIdx.x;
    if (e < rec.n) {
      float g = rec.g[e] * inv;
      const float pv = rec.p[e];
      if (wd != 0.f) g = fmaf(wd, pv, g);
      const float m = rec.m[e] * beta1 + (1.f - beta1) * g;
      const float v = rec.v[e] * beta2 + (1.f - beta2) * g * g;
      rec.m[e] = m;
      rec.v[e] = v;
      rec.p[e] = pv - step_size * (m / (sqrtf(v) + eps));
    }
  }
}

}  // namespace
}  // namespace exvae

using namespace exvae;

#define EW_LAUNCH(kernel, n, ...)                                        \
  do {                                                                   \
    kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(__VA_ARGS__);    \
    EXVAE_RETURN_LAST_ERROR();                                           \
  } while (0)
#define ROW_LAUNCH(kernel, B, ...)                                       \
  do {                                                                   \
    kernel<<<ceil_div(B, 8), 256, 0, as_stream(stream)>>>(__VA_ARGS__);  \
    EXVAE_RETURN_LAST_ERROR();                                           \
  } while (0)

extern "C" int exvae_reparameterize_fwd(const float* mu, const float* logvar, const float* eps, int64_t n, float* z,
                                        exvae_stream_t stream) {
  EXVAE_CHECK_ARG(mu && logvar && eps && z && n > 0);
  EW_LAUNCH(reparam_fwd_kernel, n, mu, logvar, eps, n, z);
}
extern "C" int exvae_reparameterize_bwd(const float* logvar, const float* eps, const float* dz, int64_t n, float* dmu,
                                        float* dlogvar, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(logvar && eps && dz && n > 0);
  EW_LAUNCH(reparam_bwd_kernel, n, logvar, eps, dz, n, dmu, dlogvar);
}
extern "C" int exvae_log_normal_diag_fwd(const float* x, const float* mean, const float* logvar, int B, int D,
                                         float* out, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(x && mean && logvar && out && B > 0 && D > 0);
  ROW_LAUNCH(lognormal_fwd_kernel, B, x, mean, logvar, B, D, out);
}
extern "C" int exvae_log_normal_diag_bwd(const float* x, const float* mean, const float* logvar, const float* dout,
                                         int B, int D, float* dx, float* dmean, float* dlogvar,
                                         exvae_stream_t stream) {
  EXVAE_CHECK_ARG(x && mean && logvar && dout && B > 0 && D > 0);
  const long long n = (long long)B * D;
  EW_LAUNCH(lognormal_bwd_kernel, n, x, mean, logvar, dout, n, D, dx, dmean, dlogvar);
}
extern "C" int exvae_log_normal_standard_fwd(const float* x, int B, int D, float* out, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(x && out && B > 0 && D > 0);
  ROW_LAUNCH(lognormstd_fwd_kernel, B, x, B, D, out);
}
extern "C" int exvae_log_normal_standard_bwd(const float* x, const float* dout, int B, int D, float* dx,
                                             exvae_stream_t stream) {
  EXVAE_CHECK_ARG(x && dout && dx && B > 0 && D > 0);
  const long long n = (long long)B * D;
  EW_LAUNCH(lognormstd_bwd_kernel, n, x, dout, n, D, dx);
}
extern "C" int exvae_log_bernoulli_fwd(const float* x, const float* mean, int B, int P, float* out,
                                       exvae_stream_t stream) {
  EXVAE_CHECK_ARG(x && mean && out && B > 0 && P > 0);
  bernoulli_fwd_kernel<<<B, 128, 0, as_stream(stream)>>>(x, mean, B, P, out);
  EXVAE_RETURN_LAST_ERROR();
}
extern "C" int exvae_log_bernoulli_bwd(const float* x, const float* mean, const float* dout, int B, int P,
                                       float* dmean, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(x && mean && dout && dmean && B > 0 && P > 0);
  const long long n = (long long)B * P;
  EW_LAUNCH(bernoulli_bwd_kernel, n, x, mean, dout, n, P, dmean);
}
extern "C" int exvae_log_logistic256_fwd(const float* x, const float* mean, const float* logvar, int B, int P,
                                         float* out, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(x && mean && logvar && out && B > 0 && P > 0);
  ROW_LAUNCH(logistic_fwd_kernel, B, x, mean, logvar, B, P, out);
}
extern "C" int exvae_log_logistic256_bwd(const float* x, const float* mean, const float* logvar, const float* dout,
                                         int B, int P, float* dmean, float* dlogvar, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(x && mean && logvar && dout && B > 0 && P > 0);
  const long long n = (long long)B * P;
  EW_LAUNCH(logistic_bwd_kernel, n, x, mean, logvar, dout, n, P, dmean, dlogvar);
}
extern "C" int exvae_elbo_reduce(const float* RE, const float* KL, int B, float beta, const float* beta_dev, int average,
                                 float* out3, float* loss_b, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(RE && KL && B > 0);
  EXVAE_CHECK_ARG(average ? out3 != nullptr : loss_b != nullptr);
  // 256 threads for a training batch: a 1024-thread CTA does not fit next to a resident 320-thread / 128-register CTA of
  // the exemplar-prior backward (register file), and then waits for a whole SM to drain (measured: 17 us on the step's
  // critical path for a 2 us reduction)
  elbo_reduce_kernel<<<1, B <= 8192 ? 256 : 1024, 0, as_stream(stream)>>>(RE, KL, B, beta, beta_dev, average, out3, loss_b);
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_lincomb4(const float* x0, const float* x1, const float* x2, const float* x3, float c0, float c1,
                              float c2, float c3, int64_t n, float* out, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(out && n > 0 && (x0 || x1 || x2 || x3));
  EW_LAUNCH(lincomb4_kernel, n, x0, x1, x2, x3, c0, c1, c2, c3, n, out);
}

extern "C" int exvae_elbo_reduce_bwd(const float* g3, const float* g_loss_b, int B, float beta, const float* beta_dev,
                                     int average, float* dRE, float* dKL, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(dRE && dKL && B > 0);
  EXVAE_CHECK_ARG(average ? g3 != nullptr : g_loss_b != nullptr);
  elbo_reduce_bwd_kernel<<<ceil_div(B, 256), 256, 0, as_stream(stream)>>>(g3, g_loss_b, B, beta, beta_dev, average, dRE, dKL);
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_rng_bernoulli(const float* p, int64_t n, uint64_t seed, uint64_t* counter, uint64_t subseq,
                                   int advance, float* out, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(p && out && n > 0);
  EW_LAUNCH(rng_bernoulli_kernel, (n + 3) / 4, p, n, seed, counter, subseq, advance, out);
}
extern "C" int exvae_rng_normal(int64_t n, uint64_t seed, uint64_t* counter, uint64_t subseq, int advance, float* out,
                                exvae_stream_t stream) {
  EXVAE_CHECK_ARG(out && n > 0);
  EW_LAUNCH(rng_normal_kernel, (n + 3) / 4, n, seed, counter, subseq, advance, out);
}
extern "C" int exvae_rng_randint(int64_t low, int64_t high, int64_t n, uint64_t seed, uint64_t* counter,
                                 uint64_t subseq, int advance, int64_t* out, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(out && n > 0 && high > low);
  EW_LAUNCH(rng_randint_kernel, (n + 1) / 2, low, (unsigned long long)(high - low), n, seed, counter, subseq, advance,
            out);
}
extern "C" int exvae_rng_advance(uint64_t* counter, uint64_t by, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(counter != nullptr);
  rng_advance_kernel<<<1, 1, 0, as_stream(stream)>>>(counter, by);
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_adam_normgrad_step(const int64_t* table, int n_tensors, int64_t max_numel, float lr, float beta1,
                                        float beta2, float eps, float weight_decay, int64_t* step, float* norms,
                                        exvae_stream_t stream) {
  EXVAE_CHECK_ARG(table && step && norms && n_tensors > 0 && max_numel > 0);
  static_assert(sizeof(AdamRec) == 5 * sizeof(int64_t), "table record is 5 x int64");
  cudaStream_t st = as_stream(stream);
  const AdamRec* recs = reinterpret_cast<const AdamRec*>(table);
  float* step_size = norms + (size_t)n_tensors * ADAM_NCH;      // one extra float behind the partial norms
  adam_norm_kernel<<<dim3(n_tensors, ADAM_NCH), 256, 0, st>>>(recs, norms, step, lr, beta1, beta2, step_size);
  EXVAE_CUDA(cudaGetLastError());
  dim3 grid((unsigned)((max_numel + 1023) / 1024), n_tensors);
  adam_update_kernel<<<grid, 256, 0, st>>>(recs, norms, step_size, beta1, beta2, eps, weight_decay);
  EXVAE_RETURN_LAST_ERROR();
}
