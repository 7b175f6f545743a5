// VampPrior (prior == 'vampprior', models/BaseModel.py:84-96,111-128): a mixture of C (~500) diagonal Gaussians whose
// means AND log-variances both come from encoding learned pseudo-inputs, so the single-contraction form of the exemplar
// prior (one shared log-variance) does not apply.  C is small: the [B,C] log-density matrix (1 MB at B=512) is
// materialised, exactly like the reference does, and the backward recomputes the weights from it.
//
//   logit[b,c] = sum_d -0.5*(lv[c,d] + log(2pi) + (z[b,d]-m[c,d])^2 / exp(lv[c,d])) - log C     (utils/distributions.py:28-33)
//   log_p[b]   = max_c logit + log sum_c exp(logit - max)                                         (BaseModel.py:123-125)
#include "common.cuh"

namespace exvae {
namespace {

constexpr float kLog2Pi = 1.8378770664093453f;

// thread = one (b, c) pair; the z row sits in shared memory, a component's row is read once per thread
__global__ void __launch_bounds__(256) vamp_logit_kernel(const float* __restrict__ z, const float* __restrict__ mean,
                                                         const float* __restrict__ logvar, int B, int C, int D,
                                                         float log_c, float* __restrict__ out) {
  extern __shared__ float zrow[];
  const int b = blockIdx.y;
  for (int d = threadIdx.x; d < D; d += blockDim.x) zrow[d] = z[(size_t)b * D + d];
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float* m = mean + (size_t)c * D;
  const float* lv = logvar + (size_t)c * D;
  float acc = 0.f;
  for (int d = 0; d < D; ++d) {
    const float diff = zrow[d] - m[d];
    acc += -0.5f * (lv[d] + kLog2Pi + diff * diff / expf(lv[d]));
  }
  out[(size_t)b * C + c] = acc - log_c;
}

// one warp per row: max-shifted log-sum-exp
__global__ void __launch_bounds__(256) vamp_row_lse_kernel(const float* __restrict__ mat, int B, int C,
                                                           float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* row = mat + (size_t)b * C;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, row[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(row[c] - mx);
  s = warp_sum(s);
  if (lane == 0) out[b] = mx + logf(s);
}

// dz[b,d] = sum_c W[b,c] * -(z[b,d]-m[c,d]) / exp(lv[c,d]),   W[b,c] = g_b * exp(logit[b,c] - log_p[b]);  thread = (b, d)
__global__ void __launch_bounds__(256) vamp_bwd_z_kernel(const float* __restrict__ z, const float* __restrict__ mean,
                                                         const float* __restrict__ logvar, const float* __restrict__ mat,
                                                         const float* __restrict__ log_p, const float* __restrict__ g,
                                                         int B, int C, int D, float* __restrict__ dz) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= (long long)B * D) return;
  const int b = (int)(e / D), d = (int)(e - (long long)b * D);
  const float zb = z[e], lp = log_p[b], gb = g[b];
  const float* row = mat + (size_t)b * C;
  float acc = 0.f;
  for (int c = 0; c < C; ++c) {
    const float w = gb * expf(row[c] - lp);
    acc -= w * (zb - mean[(size_t)c * D + d]) / expf(logvar[(size_t)c * D + d]);
  }
  dz[e] = acc;
}

// dmean[c,d] = sum_b W (z-m)/exp(lv);  dlogvar[c,d] = sum_b W * 0.5*((z-m)^2/exp(lv) - 1);  thread = (c, d)
__global__ void __launch_bounds__(256) vamp_bwd_bank_kernel(const float* __restrict__ z, const float* __restrict__ mean,
                                                            const float* __restrict__ logvar,
                                                            const float* __restrict__ mat, const float* __restrict__ log_p,
                                                            const float* __restrict__ g, int B, int C, int D,
                                                            float* __restrict__ dmean, float* __restrict__ dlogvar) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= (long long)C * D) return;
  const int c = (int)(e / D), d = (int)(e - (long long)c * D);
  const float m = mean[e], iv = 1.f / expf(logvar[e]);
  float am = 0.f, av = 0.f;
  for (int b = 0; b < B; ++b) {
    const float w = g[b] * expf(mat[(size_t)b * C + c] - log_p[b]);
    const float diff = z[(size_t)b * D + d] - m;
    am += w * diff * iv;
    av += w * 0.5f * (diff * diff * iv - 1.f);
  }
  dmean[e] = am;
  dlogvar[e] = av;
}

}  // namespace
}  // namespace exvae

using namespace exvae;

extern "C" int exvae_vamp_logprob_matrix(const float* z, const float* mean, const float* logvar, int B, int C, int D,
                                         float* out, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(z && mean && logvar && out && B > 0 && C > 0 && D > 0);
  dim3 grid(ceil_div(C, 256), B);
  vamp_logit_kernel<<<grid, 256, sizeof(float) * D, as_stream(stream)>>>(z, mean, logvar, B, C, D, logf((float)C), out);
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_vamp_lse_fwd(const float* z, const float* mean, const float* logvar, int B, int C, int D, float* mat,
                                  float* log_p, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(mat && log_p);
  int rc = exvae_vamp_logprob_matrix(z, mean, logvar, B, C, D, mat, stream);
  if (rc) return rc;
  vamp_row_lse_kernel<<<ceil_div(B, 8), 256, 0, as_stream(stream)>>>(mat, B, C, log_p);
  EXVAE_RETURN_LAST_ERROR();
}

extern "C" int exvae_vamp_lse_bwd(const float* z, const float* mean, const float* logvar, const float* mat,
                                  const float* log_p, const float* grad_log_p, int B, int C, int D, float* dz,
                                  float* dmean, float* dlogvar, exvae_stream_t stream) {
  EXVAE_CHECK_ARG(z && mean && logvar && mat && log_p && grad_log_p && dz && dmean && dlogvar);
  EXVAE_CHECK_ARG(B > 0 && C > 0 && D > 0);
  cudaStream_t st = as_stream(stream);
  vamp_bwd_z_kernel<<<ceil_div((long long)B * D, 256), 256, 0, st>>>(z, mean, logvar, mat, log_p, grad_log_p, B, C, D, dz);
  EXVAE_CUDA(cudaGetLastError());
  vamp_bwd_bank_kernel<<<ceil_div((long long)C * D, 256), 256, 0, st>>>(z, mean, logvar, mat, log_p, grad_log_p, B, C, D,
                                                                        dmean, dlogvar);
  EXVAE_RETURN_LAST_ERROR();
}
