// K1 forward as ONE kernel (sm_100a, D <= 63): pairwise squared distance -> per-pair Gaussian log-density ->
// leave-one-out mask -> log-sum-exp over the exemplars -> normaliser, for
//     log p(z_b) = LSE_n log N(z_b | mu_n, sigma^2) - log(C - #masked_b)
// (utils/distributions.py:12-25, models/BaseModel.py:98-128) straight from the RAW inputs z [B,D], mu [C,D],
// logvar [D], z_idx [B], mu_idx [C].  Nothing is staged in HBM: the former stage / mask-list / merge / finalize launches
// (and their 14 MB of staged operand planes) are folded into the tcgen05 kernel.
//
// CTA = one block of 128 latents x a range of 128-exemplar bank tiles (grid = column splits x row blocks ~ 148 CTAs),
// 544 threads:
//   all       prologue: the z tile is scaled by 1/sigma (and log2 e), augmented (z' = (zs*log2e | 1), so that
//             z'.m' is the base-2 logit), split into tf32 hi/lo and written into the SWIZZLE_128B K-major A operand;
//             the 128 dataset indices of the rows go into a 256-slot shared-memory hash set
//   warp 0    MMA issuer: 3 x KP/8 tcgen05.mma.kind::tf32 per tile into one of two TMEM accumulators
//   warps 1-8 epilogue: thread = (row, 64-column half); tcgen05.ld -> mask -> online base-2 max / sum
//   warps 9-16 converters: two threads per exemplar of the tile: raw mu row -> m' = (mu/sigma | -0.5|mu/sigma|^2 log2e),
//             hi/lo split, swizzled B operand stage (2-stage ring); probes the hash set with the exemplar's dataset
//             index and lists the (rare) columns that can be masked for this row block
// The last CTA of every row block (atomic ticket) merges the per-split (max, sum, count) partials and writes log p(z)
// (and the base-2 row log-sum for the backward), or the per-shard statistics when the bank is range-sharded.
#include "prior_lse_tc.cuh"

#include <algorithm>

#include "gemm_tc.cuh"
#include "tc_common.cuh"

namespace exvae {
using namespace tc;
namespace {

constexpr int PF_THREADS = 544;
constexpr int PF_EPI_WARPS = 8;
constexpr int PF_TILE = 128 * 128;                  // one [128 rows x 32 floats] swizzled box
constexpr long long kPadKey = INT64_MIN;
constexpr int PF_HASH = 256;
constexpr int PF_HITS = 128;                        // columns of one tile that can be masked (<= 128 by construction)

__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// 3xTF32 operand split in 5 integer / fp32 instructions per element (cvt.rna.tf32.f32 is emulated with ~4 instructions
// each on sm_100a, and the converter warps are latency bound): adding half an ulp of the 10-bit mantissa to the bit
// pattern and clearing the 13 low bits is round-to-nearest (ties away), i.e. what cvt.rna.tf32.f32 returns for finite
// values.  hi = rna(x); x - hi is exact; lo = rna(x - hi).  BOTH parts are rounded, not truncated: the logits are
// differences of terms ~|mu/sigma|^2 (1e3), and a truncated hi makes |lo| twice as large and the dropped lo*lo term
// four times as large (measured: 5e-4 absolute on log p(z), outside the 5e-5 relative bar of the goldens).
__device__ __forceinline__ void split_tf32(float x, uint32_t& h, uint32_t& l) {
  h = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  const float r = x - __uint_as_float(h);
  l = (__float_as_uint(r) + 0x1000u) & 0xffffe000u;
}
// byte offset of element (row r, k) inside the [nkb][2 planes] operand region: SWIZZLE_128B K-major tiles of 32 floats
__device__ __forceinline__ uint32_t op_off(int r, int k, int plane) {
  const int kb = k >> 5, kk = k & 31;
  return (uint32_t)((kb * 2 + plane) * PF_TILE + r * 128 + ((((kk >> 2) ^ (r & 7))) << 4) + (kk & 3) * 4);
}
__device__ __forceinline__ void put_split(unsigned char* base, int r, int k, float x) {
  uint32_t h, l;
  split_tf32(x, h, l);
  *reinterpret_cast<uint32_t*>(base + op_off(r, k, 0)) = h;
  *reinterpret_cast<uint32_t*>(base + op_off(r, k, 1)) = l;
}
__device__ __forceinline__ unsigned long long pf_gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// debug (trace runs only): cycles spent inside a barrier wait
template <typename F>
__device__ __forceinline__ void pf_timed_wait(bool on, long long& acc, F&& wait) {
  if (on) {
    const long long t0 = clock64();
    wait();
    acc += clock64() - t0;
  } else {
    wait();
  }
}
__device__ __forceinline__ uint32_t hash_key(long long k) {
  unsigned long long x = (unsigned long long)k * 0x9E3779B97F4A7C15ull;
  return (uint32_t)(x >> 56);                        // 8 bits
}

struct PfParams {
  const float* z; const float* mu; const float* logvar;
  const int64_t* z_idx; const int64_t* mu_idx;
  const int* c_valid;
  int B, C, D, KP, ntile, nsplit;
  float c_total;
  float* part;                 // [Bpad][nsplit][4]
  unsigned int* tickets;       // [row blocks], zero on entry, self-resetting
  float* stats;                // [B][4] or null
  float* log_p; float* lse2;   // [B] or null
  // by-product for the backward (null = forward only): the staged operands prior_stage_kernel would have written
  float* st_zs; float* st_ms;  // [Bpad][LD], [Cpad][LD]  z / sigma, mu / sigma (zero padded)
  float* st_zp; float* st_mp;  // [2][Bpad][KP], [2][Cpad][KP] hi / lo planes of the augmented rows
  int64_t* st_cidx;            // [Cpad] dataset index per exemplar (INT64_MIN beyond the valid rows)
  float* st_isig;              // [LD]
  int LD, Bpad, Cpad;
  unsigned long long* trace;   // debug (exvae_gemm_set_trace): 8 words per CTA, null in production
};

template <bool MASK, int NJ>
__global__ void __launch_bounds__(PF_THREADS, 1) prior_fused_fwd_kernel(const PfParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  unsigned char* sA = smem;                                   // 4 tiles (2 k-blocks x hi/lo)
  unsigned char* sB = smem + 4 * PF_TILE;                     // 2 stages x 4 tiles
  unsigned char* tail = smem + 12 * PF_TILE;
  float* isig = reinterpret_cast<float*>(tail);               // [64]
  long long* hkeys = reinterpret_cast<long long*>(tail + 256);                     // [256]
  int* hits_n = reinterpret_cast<int*>(tail + 256 + 2048);                         // [4]
  int* hits_c = reinterpret_cast<int*>(tail + 256 + 2048 + 16);                    // [4][128] column inside the tile
  long long* hits_k = reinterpret_cast<long long*>(tail + 256 + 2048 + 16 + 2048); // [4][128] its dataset index
  float2* red = reinterpret_cast<float2*>(tail + 256 + 2048 + 16 + 2048 + 4096);   // [128] (m, s) of column half 1
  float* redc = reinterpret_cast<float*>(red + 128);                                // [128] masked count of half 1
  uint64_t* bars = reinterpret_cast<uint64_t*>(redc + 128);
  uint64_t* b_full = bars;          // [2] converters -> MMA
  uint64_t* b_empty = bars + 2;     // [2] MMA -> converters
  uint64_t* acc_full = bars + 4;    // [2] MMA -> epilogue
  uint64_t* acc_empty = bars + 6;   // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  int* s_last = reinterpret_cast<int*>(bars + 9);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x, rb = blockIdx.y;
  const int t0 = (int)(((long long)p.ntile * split) / p.nsplit);
  const int t1 = (int)(((long long)p.ntile * (split + 1)) / p.nsplit);
  const int D = p.D, KP = p.KP;
  const int nks = KP / 8;
  const int Cv = p.c_valid ? min(p.C, max(*p.c_valid, 0)) : p.C;
#ifdef EXVAE_PF_TRACE      // per-role wait trace (tools/prior_fwd_trace.py): debug builds only
  unsigned long long* const tr = p.trace ? p.trace + 8 * (size_t)(blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
#else
  constexpr unsigned long long* tr = nullptr;
#endif
  if (tr && tid == 0) tr[0] = pf_gtimer();

  // ------------------------------------------------------------------ prologue
  // z rows of this block: issued first, so that their global-memory latency hides behind the barrier / TMEM / zeroing
  // work below (vector path: D % 4 == 0, i.e. every 16-byte chunk of the swizzled operand row comes from one float4)
  constexpr int ZI = 4;                                       // 128 rows x <= 15 float4 over 544 threads
  const int nq = D >> 2;
  const bool zvec = (D & 3) == 0 && (reinterpret_cast<uintptr_t>(p.z) & 15) == 0;
  float4 zr[ZI];
  if (zvec) {
#pragma unroll
    for (int i = 0; i < ZI; ++i) {
      const int e = tid + i * PF_THREADS;
      zr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < 128 * nq) {
        const int r = e / nq, q = e - r * nq;
        const int b = rb * 128 + r;
        if (b < p.B) zr[i] = __ldg(reinterpret_cast<const float4*>(p.z + (size_t)b * D) + q);
      }
    }
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&b_full[i], 8);
      mbar_init(&b_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], PF_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  // zero the operand tiles (padding columns / rows must be 0), clear the hash set and the hit counters
  for (int i = tid; i < 12 * PF_TILE / 16; i += PF_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < PF_HASH; i += PF_THREADS) hkeys[i] = kPadKey;
  if (tid < 4) hits_n[tid] = 0;
  if (tid < 64) isig[tid] = tid < D ? 1.0f / expf(0.5f * p.logvar[tid]) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // A operand: z' = (z / sigma * log2e | 1 | 0..), hi/lo tf32 planes
  const bool stage_z = p.st_zs != nullptr && split == 0;      // one CTA per row block also leaves the staged z rows
  if (zvec) {
#pragma unroll
    for (int i = 0; i < ZI; ++i) {
      const int e = tid + i * PF_THREADS;
      if (e < 128 * nq) {
        const int r = e / nq, q = e - r * nq;
        const int b = rb * 128 + r;
        if (b < p.B) {
          const float4 is4 = *reinterpret_cast<const float4*>(isig + 4 * q);
          const float4 zs4 = make_float4(zr[i].x * is4.x, zr[i].y * is4.y, zr[i].z * is4.z, zr[i].w * is4.w);
          const float v[4] = {zs4.x * kLog2e, zs4.y * kLog2e, zs4.z * kLog2e, zs4.w * kLog2e};
          uint32_t h[4], l[4];
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) split_tf32(v[c4], h[c4], l[c4]);
          const uint32_t off = (uint32_t)((q >> 3) * 2 * PF_TILE + r * 128 + ((((q & 7) ^ (r & 7))) << 4));
          *reinterpret_cast<uint4*>(sA + off) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(sA + off + PF_TILE) = make_uint4(l[0], l[1], l[2], l[3]);
          if (stage_z) *reinterpret_cast<float4*>(p.st_zs + (size_t)b * p.LD + 4 * q) = zs4;   // LD % 4 == 0
        }
      }
    }
  } else {
    for (int e = tid; e < 128 * D; e += PF_THREADS) {
      const int r = e / D, k = e - r * D;
      const int b = rb * 128 + r;
      if (b < p.B) {
        const float zsv = p.z[(size_t)b * D + k] * isig[k];
        put_split(sA, r, k, zsv * kLog2e);
        if (stage_z) p.st_zs[(size_t)b * p.LD + k] = zsv;
      }
    }
  }
  if (stage_z) {
    for (int e = tid; e < 128 * p.LD; e += PF_THREADS) {      // zero padding of the staged rows
      const int r = e / p.LD, k = e - r * p.LD;
      const int b = rb * 128 + r;
      if (b >= p.B || k >= D) p.st_zs[(size_t)b * p.LD + k] = 0.f;
    }
    if (rb == 0 && tid < p.LD) p.st_isig[tid] = tid < D ? isig[tid] : 0.f;
  }
  if (tid < 128) {
    const int b = rb * 128 + tid;
    if (b < p.B) {
      put_split(sA, tid, D, 1.0f);
      if (MASK) {                                             // hash SET of the row block's dataset indices
        const long long key = p.z_idx[b];
        uint32_t slot = hash_key(key);
        while (true) {
          const long long old = (long long)atomicCAS(reinterpret_cast<unsigned long long*>(&hkeys[slot]),
                                                     (unsigned long long)kPadKey, (unsigned long long)key);
          if (old == kPadKey || old == key) break;
          slot = (slot + 1) & (PF_HASH - 1);
        }
      }
    }
  }
  fence_proxy_async();
  __syncthreads();
  if (stage_z) {
    // the hi / lo planes of z' exactly as the tensor core sees them: copy the operand tile out of shared memory
    const size_t plane = (size_t)p.Bpad * KP;
    for (int e = tid; e < 128 * KP; e += PF_THREADS) {
      const int r = e / KP, k = e - r * KP;
      const size_t o = (size_t)(rb * 128 + r) * KP + k;
      p.st_zp[o] = __uint_as_float(*reinterpret_cast<const uint32_t*>(sA + op_off(r, k, 0)));
      p.st_zp[plane + o] = __uint_as_float(*reinterpret_cast<const uint32_t*>(sA + op_off(r, k, 1)));
    }
  }

  if (tr && tid == 0) tr[1] = pf_gtimer();
  if (warp == 0) {
    // ------------------------------------------------------------- MMA issuer
    if (elect_one_sync()) {
      long long w_full = 0, w_acc = 0;
      constexpr uint32_t idesc = umma_idesc(128, 128, false, false);
      const uint32_t a0 = smem_u32(sA);
      for (int t = t0; t < t1; ++t) {
        const int it = t - t0, s = it & 1, ph = (it >> 1) & 1;
        pf_timed_wait(tr != nullptr, w_full, [&] { mbar_wait(&b_full[s], ph); });
        pf_timed_wait(tr != nullptr, w_acc, [&] { mbar_wait(&acc_empty[s], ph ^ 1); });
        tc_fence_after();
        const uint32_t b0 = smem_u32(sB + s * 4 * PF_TILE);
        for (int ks = 0; ks < nks; ++ks) {
          const int kb = ks >> 2, kk = ks & 3;
          const uint32_t off_hi = (kb * 2 + 0) * PF_TILE + kk * 32, off_lo = (kb * 2 + 1) * PF_TILE + kk * 32;
          const uint64_t a_hi = umma_desc(a0 + off_hi, 16, 1024, 2), a_lo = umma_desc(a0 + off_lo, 16, 1024, 2);
          const uint64_t b_hi = umma_desc(b0 + off_hi, 16, 1024, 2), b_lo = umma_desc(b0 + off_lo, 16, 1024, 2);
          const uint32_t d = tmem_base + s * 128;
          umma_tf32(d, a_lo, b_hi, idesc, ks > 0 ? 1u : 0u);
          umma_tf32(d, a_hi, b_lo, idesc, 1u);
          umma_tf32(d, a_hi, b_hi, idesc, 1u);
        }
        umma_commit(&b_empty[s]);
        umma_commit(&acc_full[s]);
      }
      if (tr) { tr[3] = (unsigned long long)w_full; tr[4] = (unsigned long long)w_acc; }
    }
  } else if (warp <= PF_EPI_WARPS) {
    // ------------------------------------------------------------- epilogue warps 1..8
    const int q = warp & 3;                    // TMEM lane quadrant this warp may read
    const int half = (warp - 1) >> 2;          // which 64 of the tile's 128 columns
    const int r = 32 * q + lane;
    const int b = rb * 128 + r;
    const long long zi = (MASK && b < p.B) ? p.z_idx[b] : kPadKey;
    float m = -INFINITY, ssum = 0.f, cnt = 0.f;
    long long w_epi = 0;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * q) << 16) + half * 64;
    for (int t = t0; t < t1; ++t) {
      const int it = t - t0, s = it & 1, ph = (it >> 1) & 1;
      pf_timed_wait(tr != nullptr, w_epi, [&] { mbar_wait(&acc_full[s], ph); });
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(lane_addr + s * 128, v0);
      tmem_ld32(lane_addr + s * 128 + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      if (MASK) {
        const int ring = it & 3;
        const int nh = hits_n[ring];                          // written before b_full of this tile: visible
        for (int h = 0; h < nh; ++h) {                        // rare: columns whose index occurs in this row block
          const int jj = hits_c[ring * PF_HITS + h] - half * 64;
          if (jj >= 0 && jj < 64 && hits_k[ring * PF_HITS + h] == zi) {
            cnt += 1.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j == jj) v0[j] = 0xff800000u;
              if (j + 32 == jj) v1[j] = 0xff800000u;
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[s]);              // accumulator (and hit-list slot) drained
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fmaxf(__uint_as_float(v0[j]), __uint_as_float(v1[j])));
      const float mn = fmaxf(m, mx);
      const float ms = (mn == -INFINITY) ? 0.f : mn;
      float acc = ssum * ex2_approx(m - ms);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        acc += ex2_approx(__uint_as_float(v0[j]) - ms);
        acc += ex2_approx(__uint_as_float(v1[j]) - ms);
      }
      ssum = acc;
      m = mn;
    }
    if (tr && tid == 32) tr[5] = (unsigned long long)w_epi;
    if (half == 1) {
      red[r] = make_float2(m, ssum);
      redc[r] = cnt;
    }
    named_bar(1, PF_EPI_WARPS * 32);
    if (half == 0) {
      const float2 o = red[r];
      lse2_merge(m, ssum, o.x, o.y);
      reinterpret_cast<float4*>(p.part)[(size_t)b * p.nsplit + split] = make_float4(m, ssum, cnt + redc[r], 0.f);
    }
  } else {
    // ------------------------------------------------------------- converter warps 9..16: two threads per exemplar
    // (even / odd 16-byte chunks of its row); 4 consecutive k-values = one swizzled chunk = one 128-bit store per plane.
    // The raw rows (and dataset indices) of tile t+1 are loaded into registers BEFORE tile t is converted, so that the
    // global-memory latency hides behind the conversion and the wait for the stage (NJ = chunks per thread).
    const int ci = tid - 32 * (1 + PF_EPI_WARPS);             // 0..255
    const int c = ci >> 1, par = ci & 1;
    const int nch = (D + 4) >> 2;                             // chunks that hold data incl. the augmented column D
    const int nrb = gridDim.y;
    const bool vec = (D & 3) == 0 && (reinterpret_cast<uintptr_t>(p.mu) & 15) == 0;
    auto load_tile = [&](int t, float4 (&raw)[NJ], long long& key) {
      const int n = t * 128 + c;
      const bool valid = n < Cv;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int ch = 2 * j + par;
        raw[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && 4 * ch < D) {
          const float* src = p.mu + (size_t)n * D + 4 * ch;
          if (vec) {
            raw[j] = __ldg(reinterpret_cast<const float4*>(src));
          } else {
            raw[j].x = src[0];
            if (4 * ch + 1 < D) raw[j].y = src[1];
            if (4 * ch + 2 < D) raw[j].z = src[2];
            if (4 * ch + 3 < D) raw[j].w = src[3];
          }
        }
      }
      key = kPadKey;
      if (MASK && valid && par == 0) key = p.mu_idx[n];
    };
    float4 sc[NJ];
    long long key;
    long long w_conv = 0;
    if (t0 < t1) load_tile(t0, sc, key);
    for (int t = t0; t < t1; ++t) {
      const int it = t - t0, s = it & 1, ph = (it >> 1) & 1, ring = it & 3;
      const int n = t * 128 + c;
      const bool valid = n < Cv;
      // ONE of the row blocks' CTAs that process tile t also leaves its staged rows for the backward (round-robin, so
      // that the extra stores spread over the CTAs)
      const bool stage_m = p.st_ms != nullptr && rb == (t % nrb);
      float4 nxt[NJ];
      long long nkey = kPadKey;
      if (t + 1 < t1) load_tile(t + 1, nxt, nkey);
      pf_timed_wait(tr != nullptr, w_conv, [&] { mbar_wait(&b_empty[s], ph ^ 1); });
      unsigned char* sb = sB + s * 4 * PF_TILE;
      if (ci == 0) hits_n[ring] = 0;       // slot free: the epilogue of tile it-4 released it long ago (see ring depth)
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;           // independent chains: the warps here are latency bound
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int ch = 2 * j + par;
        const float4 is4 = *reinterpret_cast<const float4*>(isig + 4 * min(ch, 15));
        sc[j].x *= is4.x; sc[j].y *= is4.y; sc[j].z *= is4.z; sc[j].w *= is4.w;
        s0 = fmaf(sc[j].x, sc[j].x, s0); s1 = fmaf(sc[j].y, sc[j].y, s1);
        s2 = fmaf(sc[j].z, sc[j].z, s2); s3 = fmaf(sc[j].w, sc[j].w, s3);
      }
      float ss = (s0 + s1) + (s2 + s3);
      ss += __shfl_xor_sync(0xffffffffu, ss, 1);             // the two threads of a row
      const float nb2 = valid ? -0.5f * ss * kLog2e : -1e30f; // finite "minus infinity": an invalid column vanishes
      named_bar(2, 256);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int ch = 2 * j + par;
        if (ch < nch) {
          float v[4] = {sc[j].x, sc[j].y, sc[j].z, sc[j].w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (4 * ch + e == D) v[e] = nb2;                  // augmented column
          uint32_t h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) split_tf32(v[e], h[e], l[e]);
          const uint32_t off = (uint32_t)((ch >> 3) * 2 * PF_TILE + c * 128 + ((((ch & 7) ^ (c & 7))) << 4));
          *reinterpret_cast<uint4*>(sb + off) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(sb + off + PF_TILE) = make_uint4(l[0], l[1], l[2], l[3]);
          if (stage_m && 4 * ch < KP) {                       // staged planes of m' for the backward (KP % 8 == 0)
            float* mp = p.st_mp + (size_t)n * KP + 4 * ch;
            *reinterpret_cast<uint4*>(mp) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(mp + (size_t)p.Cpad * KP) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
        if (stage_m && ch >= nch && 4 * ch < KP) {            // zero padding of the staged planes
          float* mp = p.st_mp + (size_t)n * KP + 4 * ch;
          *reinterpret_cast<uint4*>(mp) = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(mp + (size_t)p.Cpad * KP) = make_uint4(0u, 0u, 0u, 0u);
        }
        if (stage_m && 4 * ch < p.LD)                         // staged mu / sigma row (zero beyond D), LD % 4 == 0
          *reinterpret_cast<float4*>(p.st_ms + (size_t)n * p.LD + 4 * ch) =
              make_float4(4 * ch < D ? sc[j].x : 0.f, 4 * ch + 1 < D ? sc[j].y : 0.f, 4 * ch + 2 < D ? sc[j].z : 0.f,
                          4 * ch + 3 < D ? sc[j].w : 0.f);
      }
      if (stage_m) {
        // chunks beyond the NJ this thread converts (only when KP / LD reach past 8 * NJ floats): zero padding
        for (int ch = 2 * NJ + par; 4 * ch < max(KP, p.LD); ch += 2) {
          if (4 * ch < KP) {
            float* mp = p.st_mp + (size_t)n * KP + 4 * ch;
            *reinterpret_cast<uint4*>(mp) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(mp + (size_t)p.Cpad * KP) = make_uint4(0u, 0u, 0u, 0u);
          }
          if (4 * ch < p.LD) *reinterpret_cast<float4*>(p.st_ms + (size_t)n * p.LD + 4 * ch) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (par == 0) p.st_cidx[n] = valid ? (p.mu_idx ? p.mu_idx[n] : (int64_t)-1) : (int64_t)kPadKey;
      }
      if (MASK && valid && par == 0) {
        uint32_t slot = hash_key(key);
        while (true) {
          const long long hk = hkeys[slot];
          if (hk == kPadKey) break;
          if (hk == key) {
            const int w = atomicAdd(&hits_n[ring], 1);
            hits_c[ring * PF_HITS + w] = c;
            hits_k[ring * PF_HITS + w] = key;
            break;
          }
          slot = (slot + 1) & (PF_HASH - 1);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_full[s]);
#pragma unroll
      for (int j = 0; j < NJ; ++j) sc[j] = nxt[j];
      key = nkey;
    }
    if (tr && ci == 0) tr[2] = (unsigned long long)w_conv;
  }

  // ------------------------------------------------------------------ last CTA of the row block: merge + finalize
  tc_fence_before();
  __threadfence();
  __syncthreads();
  if (tr && tid == 0) tr[6] = pf_gtimer();
  if (tid == 0) {
    const unsigned int prev = atomicAdd(&p.tickets[rb], 1u);
    const int last = prev == (unsigned int)(p.nsplit - 1);
    if (last) p.tickets[rb] = 0u;
    *s_last = last;
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
  if (!*s_last) {
    if (tr && tid == 0) tr[7] = pf_gtimer();
    return;
  }
  __threadfence();
  // Merge of the row block's per-split partials + normaliser: FOUR threads per row (the partials of a row are contiguous,
  // so the four read interleaved 16-byte entries), every thread's loads issued in independent batches of 8 -- a single
  // thread per row walking its 37 partials one L2 round trip at a time cost 16 us at cfg2.
  float* cst_s = reinterpret_cast<float*>(red);               // [1] (the epilogue's scratch is free now)
  if (warp == 0) {
    float c0 = 0.f;
    for (int d = lane; d < D; d += 32) c0 += p.logvar[d] + kLog2Pi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c0 += __shfl_xor_sync(0xffffffffu, c0, o);
    if (lane == 0) *cst_s = c0;
  }
  __syncthreads();
  if (tid < 512) {
    const int r = tid >> 2, sub = tid & 3;
    const int b = rb * 128 + r;
    const bool rowok = b < p.B;
    float m = -INFINITY, s = 0.f, cnt = 0.f;
    const float4* pp = reinterpret_cast<const float4*>(p.part) + (size_t)b * p.nsplit;     // b < Bpad: in range
    for (int q0 = sub; q0 < p.nsplit; q0 += 32) {
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int q = q0 + 4 * i;
        v[i] = (rowok && q < p.nsplit) ? __ldcg(pp + q) : make_float4(-INFINITY, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        lse2_merge(m, s, v[i].x, v[i].y);
        cnt += v[i].z;
      }
    }
    // |z_b / sigma|^2: the four threads take every fourth 16-byte chunk (or every fourth element) of the row
    float hz = 0.f;
    if (rowok && p.log_p) {
      if (zvec) {
        for (int q = sub; q < nq; q += 4) {
          const float4 zv = __ldg(reinterpret_cast<const float4*>(p.z + (size_t)b * D) + q);
          const float4 is4 = *reinterpret_cast<const float4*>(isig + 4 * q);
          const float v0 = zv.x * is4.x, v1 = zv.y * is4.y, v2 = zv.z * is4.z, v3 = zv.w * is4.w;
          hz = fmaf(v0, v0, hz); hz = fmaf(v1, v1, hz); hz = fmaf(v2, v2, hz); hz = fmaf(v3, v3, hz);
        }
      } else {
        for (int d = sub; d < D; d += 4) {
          const float v = p.z[(size_t)b * D + d] * isig[d];
          hz = fmaf(v, v, hz);
        }
      }
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {                        // the four threads of a row are neighbouring lanes
      const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, s, o);
      lse2_merge(m, s, om, os);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      hz += __shfl_xor_sync(0xffffffffu, hz, o);
    }
    if (rowok && sub == 0) {
      if (p.stats) reinterpret_cast<float4*>(p.stats)[b] = make_float4(m, s, cnt, 0.f);
      if (p.log_p) {
        const float l2 = m + log2f(s);
        const float ct = p.c_valid ? (float)*p.c_valid : p.c_total;
        p.lse2[b] = l2;
        p.log_p[b] = (-0.5f * *cst_s - 0.5f * hz) + kLn2 * l2 - logf(ct - cnt);
      }
    }
  }
  __syncthreads();
  if (tr && tid == 0) tr[7] = pf_gtimer() | (1ull << 63);     // top bit: this CTA did the merge
}

}  // namespace

bool prior_fused_ok(int D) {
  static const bool off = [] { const char* e = getenv("EXVAE_PRIOR_FUSED"); return e && strcmp(e, "0") == 0; }();
  return !off && D + 1 <= 64 && prior_tc_enabled();
}

size_t prior_fused_ws_bytes(int B, int C) {
  (void)C;
  const int Bpad = ceil_div(B, 128) * 128;
  return 256 + sizeof(float) * 4 * (size_t)Bpad * (size_t)(2 * sm_count());
}

// ws: [tickets (256 B)][part]
int prior_fused_fwd_launch(const float* z, const float* mu, const float* logvar, const int64_t* z_idx, const int64_t* mu_idx,
                           const int* c_valid, int B, int C, int D, float c_total, float* stats, float* log_p, float* lse2,
                           void* ws, const PriorFusedStage* sg, cudaStream_t st) {
  const int Bpad = ceil_div(B, 128) * 128, rbs = Bpad / 128;
  if (rbs > 64) return EXVAE_ERR_UNSUPPORTED;
  const int ntile = ceil_div(C, 128);
  int nsplit = std::max(1, sm_count() / rbs);
  nsplit = std::min(nsplit, ntile);
  PfParams p{};
  p.z = z; p.mu = mu; p.logvar = logvar; p.z_idx = z_idx; p.mu_idx = mu_idx; p.c_valid = c_valid;
  p.B = B; p.C = C; p.D = D; p.KP = ceil_div(D + 1, 8) * 8; p.ntile = ntile; p.nsplit = nsplit; p.c_total = c_total;
  p.tickets = static_cast<unsigned int*>(ws);
  p.part = reinterpret_cast<float*>(static_cast<char*>(ws) + 256);
  p.stats = stats; p.log_p = log_p; p.lse2 = lse2;
  if (sg) {
    p.st_zs = sg->zs; p.st_ms = sg->ms; p.st_zp = sg->zp; p.st_mp = sg->mp; p.st_cidx = sg->cidx; p.st_isig = sg->isig;
    p.LD = sg->LD; p.Bpad = sg->Bpad; p.Cpad = sg->Cpad;
  }
  p.trace = tc_take_trace(8 * 400);
  EXVAE_CUDA(cudaMemsetAsync(p.tickets, 0, 256, st));          // caller-owned workspace may be uninitialised
  constexpr int SMEM = 12 * PF_TILE + 256 + 2048 + 16 + 2048 + 4096 + 1024 + 512 + 128 + 1024;
  dim3 grid(nsplit, rbs);
  const bool mask = z_idx && mu_idx;
  auto launch = [&](auto kern) -> int {
    EXVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    kern<<<grid, PF_THREADS, SMEM, st>>>(p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? EXVAE_OK : (int)e;
  };
  // NJ = 16-byte chunks of an exemplar row per converter thread (two threads per row): ceil(((D + 4) / 4) / 2)
  const int nj = (((D + 4) >> 2) + 1) >> 1;
  if (nj <= 4) return mask ? launch(prior_fused_fwd_kernel<true, 4>) : launch(prior_fused_fwd_kernel<false, 4>);
  if (nj <= 6) return mask ? launch(prior_fused_fwd_kernel<true, 6>) : launch(prior_fused_fwd_kernel<false, 6>);
  return mask ? launch(prior_fused_fwd_kernel<true, 8>) : launch(prior_fused_fwd_kernel<false, 8>);
}

}  // namespace exvae
