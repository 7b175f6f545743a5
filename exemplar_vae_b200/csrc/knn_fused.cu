// K2 — fused kNN exemplar selection (sm_100a): distance tiles + per-row running top-k in one kernel, then a
// one-warp-per-row merge of the per-split lists.
//
//   pairwise_distance(z, sub_cache).topk(k, largest=False)      models/BaseModel.py:263-264   (metric 0)
//   ((z[:,None]-mu[None])**2).sum(2)**0.5 .topk(k=20)           utils/knn_on_latent.py:4-9    (metric 1)
//
// The [B,N] distance matrix never reaches HBM.  The grid is (column splits, blocks of 64 rows) with enough splits
// to fill the 148 SMs (B = 100 rows alone would be 2 CTAs).  Every CTA walks 64x64 distance tiles of its column
// range; each tile value is produced exactly like the materialising kernel (pairdist_knn.cu): fp64 FMA accumulation
// of exact fp32 products, combined in the reference's operation order, ONE rounding to fp32 (metric 0), or the
// reference's fp32 direct-difference sum + sqrt (metric 1).  Values that beat the row's current k-th best go to a
// per-row candidate list in shared memory and a warp per row inserts them into the sorted (distance, position) list
// (ties -> lowest position).  knn_fused_merge_kernel then merges the [nsplit, B, k] lists with one warp per row
// spread over many CTAs (merging inside the tile kernel by the last CTA of a row block was measured 10x slower: 64
// rows x nsplit lists on ONE SM).  No host sync, deterministic, graph-capturable.
#include "common.cuh"

namespace exvae {
namespace {

constexpr int KF_T = 64;        // tile edge
constexpr int KF_K = 16;        // k chunk of the distance accumulation
constexpr int KF_P = KF_T + 1;
constexpr int KF_MAXK = 32;     // k <= 32: one list entry per lane

__device__ __forceinline__ bool kf_less(float v, int i, float bv, int bi) { return v < bv || (v == bv && i < bi); }

// sorted insert of (v, i) into a warp-held ascending list (lane l = entry l, entries >= k ignored)
__device__ __forceinline__ void kf_insert(float& lv, int& li, float v, int i, int lane, int k) {
  // position = number of entries strictly smaller than the candidate
  const bool smaller = lane < k && kf_less(lv, li, v, i);
  const unsigned m = __ballot_sync(0xffffffffu, smaller);
  const int pos = __popc(m);
  if (pos >= k) return;     // warp-uniform
  const float uv = __shfl_up_sync(0xffffffffu, lv, 1);
  const int ui = __shfl_up_sync(0xffffffffu, li, 1);
  if (lane == pos) {
    lv = v;
    li = i;
  } else if (lane > pos) {
    lv = uv;
    li = ui;
  }
}

struct KfSmem {
  double As[KF_K][KF_P];
  double Bs[KF_K][KF_P];
  double na[KF_T], nb[KF_T];
  float best_v[KF_T][KF_MAXK];
  int best_i[KF_T][KF_MAXK];
  float cand_v[KF_T][KF_T];
  int cand_i[KF_T][KF_T];
  int cand_n[KF_T];
  float thr_v[KF_T];      // current k-th best per row (+inf until the list is full)
  int thr_i[KF_T];
};

template <int METRIC>
__global__ void __launch_bounds__(256) knn_fused_kernel(const float* __restrict__ z, const float* __restrict__ bank, int B,
                                                        int C, int D, int k, long long pos_offset, int nsplit,
                                                        float* __restrict__ part_v, int64_t* __restrict__ part_i) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  KfSmem& S = *reinterpret_cast<KfSmem*>(smem_raw);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;
  const int split = blockIdx.x, rb = blockIdx.y;
  const int m0 = rb * KF_T;
  const int ntile = ceil_div(C, KF_T);
  const int t0 = (int)(((long long)ntile * split) / nsplit), t1 = (int)(((long long)ntile * (split + 1)) / nsplit);

  for (int e = tid; e < KF_T * KF_MAXK; e += 256) {
    (&S.best_v[0][0])[e] = INFINITY;
    (&S.best_i[0][0])[e] = 0x7fffffff;
  }
  if (tid < KF_T) {
    S.cand_n[tid] = 0;
    S.thr_v[tid] = INFINITY;
    S.thr_i[tid] = 0x7fffffff;
  }
  __syncthreads();

  for (int t = t0; t < t1; ++t) {
    const int n0 = t * KF_T;
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
    for (int k0 = 0; k0 < D; k0 += KF_K) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = tid + 256 * e;
        const int r = idx >> 4, kk = idx & 15;
        const int d = k0 + kk;
        float a = 0.f, b = 0.f;
        if (d < D && m0 + r < B) a = z[(size_t)(m0 + r) * D + d];
        if (d < D && n0 + r < C) b = bank[(size_t)(n0 + r) * D + d];
        S.As[kk][r] = (double)a;
        S.Bs[kk][r] = (double)b;
      }
      __syncthreads();
      if (METRIC == 0) {
        if (tid < KF_T) {
#pragma unroll
          for (int kk = 0; kk < KF_K; ++kk) nrm = fma(S.As[kk][tid], S.As[kk][tid], nrm);
        } else if (tid < 2 * KF_T) {
#pragma unroll
          for (int kk = 0; kk < KF_K; ++kk) nrm = fma(S.Bs[kk][tid - KF_T], S.Bs[kk][tid - KF_T], nrm);
        }
      }
#pragma unroll
      for (int kk = 0; kk < KF_K; ++kk) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = S.As[kk][ty + 16 * i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = S.Bs[kk][tx + 16 * j];
        if (METRIC == 1) {
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
    if (tid < KF_T) S.na[tid] = nrm;
    else if (tid < 2 * KF_T) S.nb[tid - KF_T] = nrm;
    __syncthreads();

    // ---- candidates: values that beat the row's current k-th best
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 16 * i;
      if (m0 + r >= B) continue;
      const float tv = S.thr_v[r];
      const int ti = S.thr_i[r];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cc = tx + 16 * j, c = n0 + cc;
        if (c >= C) continue;
        const float v = METRIC == 1 ? sqrtf(accf[i][j]) : (float)((S.na[r] + S.nb[cc]) + (-2.0 * acc[i][j]));
        if (kf_less(v, c, tv, ti)) {
          const int slot = atomicAdd(&S.cand_n[r], 1);
          S.cand_v[r][slot] = v;
          S.cand_i[r][slot] = c;
        }
      }
    }
    __syncthreads();
    // ---- one warp per row: insert the candidates (in position order: slot order is not deterministic, the sorted
    //      list is — insertion of a set of distinct (v, i) keys commutes)
    for (int r = warp; r < KF_T; r += 8) {
      const int n = S.cand_n[r];
      if (n == 0) continue;   // warp-uniform
      float lv = S.best_v[r][lane];
      int li = S.best_i[r][lane];
      for (int q = 0; q < n; ++q) kf_insert(lv, li, S.cand_v[r][q], S.cand_i[r][q], lane, k);
      S.best_v[r][lane] = lv;
      S.best_i[r][lane] = li;
      if (lane == k - 1) {
        S.thr_v[r] = lv;
        S.thr_i[r] = li;
      }
      if (lane == 0) S.cand_n[r] = 0;
    }
    __syncthreads();
  }

  // ---- emit this split's sorted list (global positions; -1 = fewer than k candidates in this split)
  for (int e = tid; e < KF_T * k; e += 256) {
    const int r = e / k, j = e - r * k;
    if (m0 + r < B) {
      const size_t o = ((size_t)split * B + (m0 + r)) * k + j;
      const int i = S.best_i[r][j];
      part_v[o] = S.best_v[r][j];
      part_i[o] = i == 0x7fffffff ? (int64_t)-1 : (int64_t)i + pos_offset;
    }
  }
}

// Merge the per-split lists (each sorted ascending): one warp per row, a k-way merge over the list HEADS.  Lane l owns
// the splits l, l+32, ... (<= 8 of them: nsplit <= 256) and keeps their current heads in registers; every pass takes
// the warp-wide smallest (dist, global position) head and only the owning lane loads that split's next entry.
__global__ void __launch_bounds__(256) knn_fused_merge_kernel(const int64_t* __restrict__ idx, const float* __restrict__ dist,
                                                              int G, int B, int k, int64_t* __restrict__ out_idx,
                                                              float* __restrict__ out_dist) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  float hv[8];
  long long hi[8];
  int hp[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int g = lane + 32 * q;
    hv[q] = INFINITY;
    hi[q] = INT64_MAX;
    hp[q] = 0;
    if (g < G) {
      const size_t o = ((size_t)g * B + b) * k;
      const long long i = idx[o];
      if (i >= 0) {
        hv[q] = dist[o];
        hi[q] = i;
      }
    }
  }
  for (int p = 0; p < k; ++p) {
    float bv = INFINITY;
    long long bi = INT64_MAX;
    int bq = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (hv[q] < bv || (hv[q] == bv && hi[q] < bi)) {
        bv = hv[q];
        bi = hi[q];
        bq = q;
      }
    float wv = bv;
    long long wi = bi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, wi, o);
      if (ov < wv || (ov == wv && oi < wi)) {
        wv = ov;
        wi = oi;
      }
    }
    if (lane == 0) {
      out_idx[(size_t)b * k + p] = wi == INT64_MAX ? -1 : wi;
      out_dist[(size_t)b * k + p] = wv;
    }
    if (wi != INT64_MAX && wi == bi && wv == bv) {       // this lane owns the winner (global positions are unique): advance
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q == bq) {
          const int g = lane + 32 * q;
          const int np = hp[q] + 1;
          hp[q] = np;
          hv[q] = INFINITY;
          hi[q] = INT64_MAX;
          if (np < k) {
            const size_t o = ((size_t)g * B + b) * k + np;
            const long long i = idx[o];
            if (i >= 0) {
              hv[q] = dist[o];
              hi[q] = i;
            }
          }
        }
    }
  }
}

inline int knn_splits(int B, int C) {
  const int rb = ceil_div(B, KF_T);
  const int ntile = ceil_div(C, KF_T);
  return std::max(1, std::min(std::min(ntile, 256), ceil_div(sm_count(), rb)));
}

struct KnnWs {
  size_t off_v, off_i, bytes;
};
inline KnnWs knn_ws(int B, int C, int k) {
  KnnWs w;
  const int ns = knn_splits(B, C);
  size_t off = 0;
  w.off_v = off; off += align_up(sizeof(float) * (size_t)ns * B * k, 256);
  w.off_i = off; off += align_up(sizeof(int64_t) * (size_t)ns * B * k, 256);
  w.bytes = off;
  return w;
}

}  // namespace
}  // namespace exvae

using namespace exvae;

extern "C" size_t exvae_knn_workspace_bytes(int B, int C, int D, int k) {
  (void)D;
  if (B <= 0 || C <= 0 || k <= 0) return 0;
  return knn_ws(B, C, k).bytes;
}

extern "C" int exvae_knn_topk(const float* z, const float* bank, int B, int C, int D, int k, int metric,
                              int64_t pos_offset, int64_t* out_idx, float* out_dist, void* ws, size_t ws_bytes,
                              exvae_stream_t stream) {
  EXVAE_CHECK_ARG(z && bank && out_idx && out_dist && ws && B > 0 && C > 0 && D > 0 && k > 0);
  EXVAE_CHECK_ARG(metric == 0 || metric == 1);
  if (k > KF_MAXK) return EXVAE_ERR_UNSUPPORTED;
  const KnnWs w = knn_ws(B, C, k);
  if (ws_bytes < w.bytes) return EXVAE_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  char* base = static_cast<char*>(ws);
  float* pv = reinterpret_cast<float*>(base + w.off_v);
  int64_t* pi = reinterpret_cast<int64_t*>(base + w.off_i);
  const int ns = knn_splits(B, C);
  dim3 grid(ns, ceil_div(B, KF_T));
  auto launch = [&](auto kern) -> int {
    EXVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KfSmem)));
    kern<<<grid, 256, sizeof(KfSmem), st>>>(z, bank, B, C, D, k, (long long)pos_offset, ns, pv, pi);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? EXVAE_OK : (int)e;
  };
  int rc = metric == 0 ? launch(knn_fused_kernel<0>) : launch(knn_fused_kernel<1>);
  if (rc) return rc;
  knn_fused_merge_kernel<<<ceil_div(B, 8), 256, 0, st>>>(pi, pv, ns, B, k, out_idx, out_dist);
  EXVAE_RETURN_LAST_ERROR();
}
