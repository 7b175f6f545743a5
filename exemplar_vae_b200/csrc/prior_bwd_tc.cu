// K1 backward on the tensor cores (sm_100a): dz, dmu and the pieces of dlogvar of the exemplar prior
// (models/BaseModel.py:98-128 differentiated), flash-attention style: the weights
//     W[b,n] = g_b * 2^(logit2[b,n] - lse2_b)          (0 for leave-one-out pairs)
// are recomputed from the saved row log-sum and never leave the SM.
//
//     dzs[b,:] = sum_n W[b,n] ms[n,:] - rowsum_b zs[b,:]          dms[n,:] = sum_b W[b,n] zs[b,:] - colsum_n ms[n,:]
//
// Both contractions need W as the *A* operand of a 128-row MMA (rows = TMEM lanes), once with rows = b and once
// with rows = n, so the same kernel runs twice with the operand roles swapped (TR = false / true):
//     S = X' . Y'^T        X' = lane-side rows (z' or m'), Y' = column-side rows; 6 k-steps x 3 tf32 products
//     W = epilogue(S)      TMEM -> registers -> exp2 / mask -> hi (raw bits) + lo -> TMEM (tcgen05.st)
//     G += W . YT^T        A = W from tensor memory, B = the transposed column-side tile (rows d, K-major)
// G accumulates in TMEM over all column-side tiles of the CTA (TR = false: the CTA's range of bank tiles for one block
// of 128 latents -> per-split partial of W.ms; TR = true: all row blocks for one tile of 128 exemplars -> W^T.zs
// complete, no cross-CTA reduction).  Error-compensated 3xTF32 as in the forward: products of hi/lo parts, the
// tensor core drops the 13 low mantissa bits of the raw fp32 W itself.
//
// CTA = 320 threads: warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue (thread = TMEM lane x 64-column
// half).  The column-side K-major tile is free again as soon as the S MMAs have retired, the transposed tile when the
// G MMAs have: the producer refills each right away, so both loads overlap the other phase without double buffers.
#include "prior_lse_tc.cuh"

#include <algorithm>

#include "gemm_tc.cuh"
#include "tc_common.cuh"

namespace exvae {
using namespace tc;
namespace {

constexpr int BT_THREADS = 320;
constexpr int BT_EPI_WARPS = 8;
constexpr int BT_BOX = 128 * 128;          // one [128 rows x 32 floats] swizzled box
constexpr long long kPad = INT64_MIN;
constexpr uint32_t TM_S = 0, TM_WLO = 128, TM_G = 256;   // TMEM columns: S / W_hi (in place), W_lo, G accumulator

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float split_lo(float x) {
  const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  const float r = x - h;
  return __uint_as_float((__float_as_uint(r) + 0x1000u) & 0xffffe000u);
}

struct BwdP {
  const float* glp; const float* lsp; const int64_t* zip; const int64_t* cidx;
  const float* zs; const float* ms; const float* isig;
  int Bpad, Cpad, KP, NG, LD, B, C, D, ny, nsplit;
  float* dzs_part; float* rowsum_part; float* dmu; float* coldot_part;
  float* gcol_part; float* tot_part;
  int x0;                      // TR: first bank tile of this launch (pass 2 goes out in single-wave chunks)
  unsigned long long* trace;   // debug (exvae_gemm_set_trace): 8 words per CTA, null in production
};

template <bool TR, bool MASK>
__global__ void __launch_bounds__(BT_THREADS, 1)
    prior_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmYK,
                        const __grid_constant__ CUtensorMap tmYT, const BwdP p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  // layout: X [2 kb][2 planes] 64 KB | YK [2 kb][2 planes] 64 KB | YT [4 kb][2 planes] NG x 128 B | slots | red | barriers
  const int yt_box = p.NG * 128;
  unsigned char* sX = smem;
  unsigned char* sYK = smem + 4 * BT_BOX;
  unsigned char* sYT = smem + 8 * BT_BOX;
  unsigned char* tail = sYT + 8 * 64 * 128;                       // sized for NG <= 64
  long long* cis = reinterpret_cast<long long*>(tail);            // [2][128] column-side dataset indices
  float* gcol = reinterpret_cast<float*>(tail + 2048);            // [2][128] TR: upstream gradient per column
  float* lcol = reinterpret_cast<float*>(tail + 3072);            // [2][128] TR: row log-sum per column
  float* red = reinterpret_cast<float*>(tail + 4096);             // [128] row sums of column half 1 | [8][64] coldot
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 4096 + 2560);
  uint64_t* x_full = bars;
  uint64_t* yk_full = bars + 1;
  uint64_t* yk_empty = bars + 2;
  uint64_t* yt_full = bars + 3;
  uint64_t* yt_empty = bars + 4;
  uint64_t* s_full = bars + 5;
  uint64_t* w_ready = bars + 6;
  uint64_t* g_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned long long* tr = p.trace ? p.trace + 8 * (size_t)(blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
  auto stamp = [&](int i) {
    if (tr) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      tr[i] = t;
    }
  };
  if (tid == 0) stamp(0);
  // TR = false: X = block blockIdx.y of 128 latents, Y = bank tiles [y0, y1) of split blockIdx.x (of nsplit)
  // TR = true : X = bank tile blockIdx.x,            Y = row blocks [y0, y1) of split blockIdx.y (of gridDim.y: a
  //             range-sharded bank has few exemplar tiles per rank, so the row blocks are split over CTAs as well
  //             and prior_bwd_cols_kernel adds the partials in a fixed order)
  const int xi = TR ? (int)blockIdx.x + p.x0 : (int)blockIdx.y;
  const int y0 = TR ? (int)(((long long)p.ny * blockIdx.y) / gridDim.y) : (int)(((long long)p.ny * blockIdx.x) / p.nsplit);
  const int y1 = TR ? (int)(((long long)p.ny * (blockIdx.y + 1)) / gridDim.y)
                    : (int)(((long long)p.ny * (blockIdx.x + 1)) / p.nsplit);
  const int nkb = (p.KP + 31) / 32;
  const int nks = p.KP / 8;

  if (tid == 0) {
    mbar_init(x_full, 1);
    mbar_init(yk_full, 1);
    mbar_init(yk_empty, 1);
    mbar_init(yt_full, 1);
    mbar_init(yt_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(w_ready, BT_EPI_WARPS);
    mbar_init(g_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(x_full, nkb * 2 * BT_BOX);
      for (int kb = 0; kb < nkb; ++kb)
        for (int pl = 0; pl < 2; ++pl) tma_load_3d(sX + (kb * 2 + pl) * BT_BOX, &tmX, x_full, kb * 32, xi * 128, pl);
      for (int y = y0; y < y1; ++y) {
        const int it = y - y0, ph = it & 1, slot = it & 1;
        mbar_wait(yk_empty, ph ^ 1);
        const int extra = (MASK ? 128 * 8 : 0) + (TR ? 2 * 128 * 4 : 0);
        mbar_arrive_expect_tx(yk_full, nkb * 2 * BT_BOX + extra);
        for (int kb = 0; kb < nkb; ++kb)
          for (int pl = 0; pl < 2; ++pl) tma_load_3d(sYK + (kb * 2 + pl) * BT_BOX, &tmYK, yk_full, kb * 32, y * 128, pl);
        if (MASK) bulk_g2s(cis + slot * 128, (TR ? p.zip : p.cidx) + (size_t)y * 128, 128 * 8, yk_full);
        if (TR) {
          bulk_g2s(gcol + slot * 128, p.glp + (size_t)y * 128, 128 * 4, yk_full);
          bulk_g2s(lcol + slot * 128, p.lsp + (size_t)y * 128, 128 * 4, yk_full);
        }
        mbar_wait(yt_empty, ph ^ 1);
        mbar_arrive_expect_tx(yt_full, 8 * yt_box);
        for (int kb = 0; kb < 4; ++kb)
          for (int pl = 0; pl < 2; ++pl)
            tma_load_3d(sYT + (kb * 2 + pl) * yt_box, &tmYT, yt_full, y * 128 + kb * 32, 0, pl);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one_sync()) {
      const uint32_t idesc_s = umma_idesc(128, 128, false, false);
      const uint32_t idesc_g = umma_idesc(128, p.NG, false, false);
      mbar_wait(x_full, 0);
      stamp(1);
      const uint32_t x0 = smem_u32(sX), k0 = smem_u32(sYK), t0 = smem_u32(sYT);
      for (int y = y0; y < y1; ++y) {
        const int it = y - y0, ph = it & 1;
        mbar_wait(yk_full, ph);
        tc_fence_after();
        if (it == 0) stamp(2);
        for (int ks = 0; ks < nks; ++ks) {
          const int kb = ks >> 2, kk = ks & 3;
          const uint32_t off_hi = (kb * 2 + 0) * BT_BOX + kk * 32, off_lo = (kb * 2 + 1) * BT_BOX + kk * 32;
          const uint64_t a_hi = umma_desc(x0 + off_hi, 16, 1024, 2), a_lo = umma_desc(x0 + off_lo, 16, 1024, 2);
          const uint64_t b_hi = umma_desc(k0 + off_hi, 16, 1024, 2), b_lo = umma_desc(k0 + off_lo, 16, 1024, 2);
          umma_tf32(tmem_base + TM_S, a_lo, b_hi, idesc_s, ks > 0 ? 1u : 0u);
          umma_tf32(tmem_base + TM_S, a_hi, b_lo, idesc_s, 1u);
          umma_tf32(tmem_base + TM_S, a_hi, b_hi, idesc_s, 1u);
        }
        umma_commit(yk_empty);          // the K-major column-side tile can be refilled
        umma_commit(s_full);            // S complete: the epilogue turns it into W
        mbar_wait(w_ready, ph);
        tc_fence_after();
        mbar_wait(yt_full, ph);
        tc_fence_after();
        for (int ks = 0; ks < 16; ++ks) {     // K = the 128 column-side rows
          const int kb = ks >> 2, kk = ks & 3;
          const uint64_t b_hi = umma_desc(t0 + (kb * 2 + 0) * yt_box + kk * 32, 16, 1024, 2);
          const uint64_t b_lo = umma_desc(t0 + (kb * 2 + 1) * yt_box + kk * 32, 16, 1024, 2);
          const uint32_t first = (it > 0 || ks > 0) ? 1u : 0u;
          umma_tf32_ts(tmem_base + TM_G, tmem_base + TM_WLO + 8 * ks, b_hi, idesc_g, first);
          umma_tf32_ts(tmem_base + TM_G, tmem_base + TM_S + 8 * ks, b_lo, idesc_g, 1u);
          umma_tf32_ts(tmem_base + TM_G, tmem_base + TM_S + 8 * ks, b_hi, idesc_g, 1u);
        }
        umma_commit(yt_empty);          // the transposed tile can be refilled (and W may be overwritten: in-order pipe)
      }
      umma_commit(g_full);
      stamp(3);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 2..9
    const int q = warp & 3;                    // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;          // which 64 of the tile's 128 columns
    const int r = 32 * q + lane;               // row inside the lane-side tile
    const int xrow = xi * 128 + r;
    float gi = 0.f, li = INFINITY;
    long long my_idx = kPad;
    if (!TR) {
      gi = p.glp[xrow];
      li = p.lsp[xrow];
      if (MASK) my_idx = p.zip[xrow];
    } else if (MASK) {
      my_idx = p.cidx[xrow];
    }
    float rowsum = 0.f;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * q) << 16);
    for (int y = y0; y < y1; ++y) {
      const int it = y - y0, ph = it & 1, slot = it & 1;
      mbar_wait(s_full, ph);
      tc_fence_after();
      uint32_t v[2][32];
      tmem_ld32(lane_addr + TM_S + half * 64, v[0]);
      tmem_ld32(lane_addr + TM_S + half * 64 + 32, v[1]);
      tmem_ld_wait();
      const long long* ci = cis + slot * 128 + half * 64;
      const float* gc = gcol + slot * 128 + half * 64;
      const float* lc = lcol + slot * 128 + half * 64;
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        uint32_t lo[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int c = 32 * h2 + j;
          const float gg = TR ? gc[c] : gi;
          const float ll = TR ? lc[c] : li;
          float w = gg * ex2_approx(__uint_as_float(v[h2][j]) - ll);
          if (MASK) {
            const long long cj = ci[c];
            if (cj == my_idx && cj != kPad) w = 0.f;
          }
          if (ll == -INFINITY) w = 0.f;        // fully masked row: the reference yields NaN; keep the gradients NaN-free
          rowsum += w;
          v[h2][j] = __float_as_uint(w);
          lo[j] = __float_as_uint(split_lo(w));
        }
        tmem_st32(lane_addr + TM_S + half * 64 + 32 * h2, v[h2]);       // W hi = raw bits, in place of S
        tmem_st32(lane_addr + TM_WLO + half * 64 + 32 * h2, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(w_ready);
    }
    // ---- drain G: each of the two warps of a quadrant takes 32 of the (<= 64) columns
    float* cd = red + 256;                                    // [4 quadrants][64] column dots (TR only)
    red[half * 128 + r] = rowsum;
    named_bar_sync(1, BT_EPI_WARPS * 32);
    const float tot = red[r] + red[128 + r];                  // sum of W over this CTA's columns, for lane r
    const int dbase = half * 32;
    uint32_t gv[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) gv[j] = 0u;
    if (y1 > y0) {
      mbar_wait(g_full, 0);
      tc_fence_after();
      if (tid == 64) stamp(4);
      if (dbase < p.NG) {
        tmem_ld32(lane_addr + TM_G + dbase, gv);
        tmem_ld_wait();
      }
    }
    if (!TR) {
      // per-split partial of W.ms and of the row sums; prior_bwd_rows_kernel finishes dz
      float* dst = p.dzs_part + ((size_t)blockIdx.x * p.Bpad + xrow) * p.LD;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (dbase + j < p.LD) dst[dbase + j] = (dbase + j < p.D) ? __uint_as_float(gv[j]) : 0.f;
      if (half == 0) p.rowsum_part[(size_t)blockIdx.x * p.Bpad + xrow] = tot;
    } else if (gridDim.y > 1) {
      // row blocks split over CTAs: emit this split's share of W^T.zs and of the column sums
      float* dst = p.gcol_part + ((size_t)blockIdx.y * p.Cpad + xrow) * p.NG + dbase;
      if (dbase < p.NG) {
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4)
          if (dbase + j4 < p.NG)
            *reinterpret_cast<float4*>(dst + j4) = make_float4(__uint_as_float(gv[j4]), __uint_as_float(gv[j4 + 1]),
                                                               __uint_as_float(gv[j4 + 2]), __uint_as_float(gv[j4 + 3]));
      }
      if (half == 0) p.tot_part[(size_t)blockIdx.y * p.Cpad + xrow] = tot;
    } else {
      // dmu[n,d] = (G[n,d] - colsum_n ms[n,d]) / sigma_d ;  coldot[d] = sum_n (G[n,d] - colsum_n ms[n,d]) ms[n,d]
      // (per-thread work first, all loads up front; the column dots go through a [128][65] shared-memory patch in the
      //  idle X region instead of 32 warp reductions per thread)
      float* pdm = reinterpret_cast<float*>(sX);              // X tile is dead: every S MMA has retired
      const float* msr = p.ms + (size_t)xrow * p.LD + dbase;
      float mv[32];
#pragma unroll
      for (int j4 = 0; j4 < 32; j4 += 4) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dbase + j4 < p.LD) t = *reinterpret_cast<const float4*>(msr + j4);   // LD % 4 == 0, rows 16-byte aligned
        mv[j4] = t.x; mv[j4 + 1] = t.y; mv[j4 + 2] = t.z; mv[j4 + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int d = dbase + j;
        const float dv = d < p.D ? __uint_as_float(gv[j]) - tot * mv[j] : 0.f;
        if (d < p.D && xrow < p.C) p.dmu[(size_t)xrow * p.D + d] = dv * p.isig[d];
        pdm[r * 65 + d] = dv * mv[j];
      }
      named_bar_sync(1, BT_EPI_WARPS * 32);
      const int e = tid - 64;                                  // 0..255: 4 row groups x 64 columns
      const int dcol = e & 63, rg = e >> 6;
      float acc = 0.f;
#pragma unroll 8
      for (int rr = 0; rr < 32; ++rr) acc += pdm[(rg * 32 + rr) * 65 + dcol];
      cd[rg * 64 + dcol] = acc;
      named_bar_sync(1, BT_EPI_WARPS * 32);
      if (e < p.LD)
        p.coldot_part[(size_t)xi * p.LD + e] = e < p.D ? (cd[e] + cd[64 + e]) + (cd[128 + e] + cd[192 + e]) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    stamp(5);
    if (tr) {
      unsigned int sm;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
      tr[6] = sm;
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// zsT / msT hi-lo planes [2][NG][rows_pad] (hi = the raw value: the tensor core truncates it itself) and the padded
// per-row arrays of the z side.  One thread per source row; blocks [0, Cpad/256) bank rows, the rest z rows.
__global__ void __launch_bounds__(256) prior_bwd_prep_kernel(const float* __restrict__ zs, const float* __restrict__ ms,
                                                             const float* __restrict__ g,
                                                             const float* __restrict__ lse2,
                                                             const int64_t* __restrict__ z_idx, int B, int D, int LD,
                                                             int Bpad, int Cpad, int nbank, int NG,
                                                             float* __restrict__ zsT,
                                                             float* __restrict__ msT, float* __restrict__ glp,
                                                             float* __restrict__ lsp, int64_t* __restrict__ zip) {
  const bool bank = (int)blockIdx.x < nbank;
  const int row = (bank ? blockIdx.x : blockIdx.x - nbank) * 256 + threadIdx.x;
  const int rows = bank ? Cpad : Bpad;
  if (row >= rows) return;
  const float* src = (bank ? ms : zs) + (size_t)row * LD;
  float* dst = bank ? msT : zsT;
  const size_t plane = (size_t)NG * rows;
  for (int d = 0; d < NG; ++d) {
    const float x = d < D ? src[d] : 0.f;       // padded rows of zs / ms are zero already
    dst[(size_t)d * rows + row] = x;
    dst[plane + (size_t)d * rows + row] = split_lo(x);
  }
  if (!bank) {
    glp[row] = row < B ? g[row] : 0.f;
    lsp[row] = row < B ? lse2[row] : INFINITY;   // +inf => weight exactly 0 for padded rows
    if (zip) zip[row] = (row < B && z_idx) ? z_idx[row] : kPad;
  }
}

// finish of a row-split pass 2: one CTA (512 threads: 4 per exemplar row, each a quarter of the float4 chunks) per
// tile of 128 rows adds the splits in a fixed order,
//   dmu[n,d] = (G[n,d] - colsum_n ms[n,d]) / sigma_d ;  coldot_part[tile][d] = sum_n (G[n,d] - colsum_n ms[n,d]) ms[n,d]
__global__ void __launch_bounds__(512) prior_bwd_cols_kernel(const float* __restrict__ gcol_part,
                                                             const float* __restrict__ tot_part,
                                                             const float* __restrict__ ms,
                                                             const float* __restrict__ isig, int rsplit, int C, int D,
                                                             int LD, int NG, int Cpad, float* __restrict__ dmu,
                                                             float* __restrict__ coldot_part) {
  extern __shared__ float pdm[];            // [128][NG + 1] products dms * ms
  const int tid = threadIdx.x;
  const int r = tid >> 2, q = tid & 3;
  const int n = blockIdx.x * 128 + r;
  const int P = NG + 1;
  float tot = 0.f;
  for (int s = 0; s < rsplit; ++s) tot += tot_part[(size_t)s * Cpad + n];
  for (int c4 = q; c4 < NG / 4; c4 += 4) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < rsplit; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(gcol_part + ((size_t)s * Cpad + n) * NG + 4 * c4);
      g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
    }
    const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int d = 4 * c4 + e;
      float pd = 0.f;
      if (d < D) {
        const float mv = ms[(size_t)n * LD + d];
        const float dv = gv[e] - tot * mv;
        if (n < C) dmu[(size_t)n * D + d] = dv * isig[d];
        pd = dv * mv;
      }
      pdm[r * P + d] = pd;
    }
  }
  __syncthreads();
  // column sums over the 128 rows: 4 row groups x NG columns, then the 4 partials in a fixed order
  float* cd = pdm + 128 * P;                // [4][NG]
  for (int e = tid; e < 4 * NG; e += 512) {
    const int d = e % NG, rg = e / NG;
    float a = 0.f;
    for (int rr = 0; rr < 32; ++rr) a += pdm[(rg * 32 + rr) * P + d];
    cd[rg * NG + d] = a;
  }
  __syncthreads();
  for (int d = tid; d < LD; d += 512)
    coldot_part[(size_t)blockIdx.x * LD + d] = d < D ? (cd[d] + cd[NG + d]) + (cd[2 * NG + d] + cd[3 * NG + d]) : 0.f;
}

}  // namespace

int prior_bwd_pass2_splits(int Bpad, int Cpad) {
  const int ntile = Cpad / 128, rbs = Bpad / 128;
  return std::max(1, std::min(rbs, sm_count() / std::max(ntile, 1)));
}

int prior_bwd_prep_launch(const float* zs, const float* ms, const float* g, const float* lse2, const int64_t* z_idx,
                          int B, int C, int D, int LD, int Bpad, int Cpad, int NG, float* zsT, float* msT, float* glp,
                          float* lsp, int64_t* zip, cudaStream_t st) {
  (void)C;
  const int nbank = ceil_div(Cpad, 256);
  prior_bwd_prep_kernel<<<nbank + ceil_div(Bpad, 256), 256, 0, st>>>(zs, ms, g, lse2, z_idx, B, D, LD, Bpad, Cpad, nbank,
                                                                    NG, zsT, msT, glp, lsp, zip);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? EXVAE_OK : (int)e;
}

int prior_bwd_tc_launch(const PriorBwdTcArgs& a, int* nsplit_out, int* ntile_out, cudaStream_t st) {
  CUtensorMap mz, mm, mzt, mmt;
  int rc = make_map(&mz, a.zp, a.Bpad, a.KP, 128, false);
  if (rc) return rc;
  rc = make_map(&mm, a.mp, a.Cpad, a.KP, 128, false);
  if (rc) return rc;
  rc = make_map(&mzt, a.zsT, a.NG, a.Bpad, a.NG, false);
  if (rc) return rc;
  rc = make_map(&mmt, a.msT, a.NG, a.Cpad, a.NG, false);
  if (rc) return rc;
  const int ntile = a.Cpad / 128, rbs = a.Bpad / 128;
  int nsplit = std::max(1, sm_count() / rbs);
  nsplit = std::min(nsplit, ntile);
  *nsplit_out = nsplit;
  *ntile_out = ntile;
  constexpr int SMEM = 8 * BT_BOX + 8 * 64 * 128 + 4096 + 2560 + 256 + 1024;
  BwdP p{};
  p.glp = a.glp; p.lsp = a.lsp; p.zip = a.zip; p.cidx = a.cidx; p.zs = a.zs; p.ms = a.ms; p.isig = a.isig;
  p.Bpad = a.Bpad; p.Cpad = a.Cpad; p.KP = a.KP; p.NG = a.NG; p.LD = a.LD; p.B = a.B; p.C = a.C; p.D = a.D;
  p.dzs_part = a.dzs_part; p.rowsum_part = a.rowsum_part; p.dmu = a.dmu; p.coldot_part = a.coldot_part;
  p.gcol_part = a.gcol_part; p.tot_part = a.tot_part;
  auto launch = [&](auto kern, dim3 grid, const CUtensorMap& x, const CUtensorMap& yk, const CUtensorMap& yt) -> int {
    EXVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    kern<<<grid, BT_THREADS, SMEM, st>>>(x, yk, yt, p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? EXVAE_OK : (int)e;
  };
  // pass 1: lanes = latents, columns = the split's bank tiles      -> dzs / rowsum partials
  p.ny = ntile; p.nsplit = nsplit;
  p.trace = tc_take_trace(8 * 400);
  rc = a.zip ? launch(prior_bwd_tc_kernel<false, true>, dim3(nsplit, rbs), mz, mm, mmt)
             : launch(prior_bwd_tc_kernel<false, false>, dim3(nsplit, rbs), mz, mm, mmt);
  if (rc) return rc;
  // pass 2: lanes = exemplars of one tile, columns = all row blocks -> dmu, coldot
  p.ny = rbs; p.nsplit = 1;
  p.trace = tc_take_trace(8 * 400);
  const int rsplit = (a.gcol_part && a.tot_part) ? prior_bwd_pass2_splits(a.Bpad, a.Cpad) : 1;
  // More tiles than SMs (cfg2: 196): equal single-wave chunks (2 x 98) instead of one grid of 148 + 48 CTAs.  The time is
  // the same two waves, but no launch ever has CTAs PENDING -- the block scheduler dispatches in order, so a kernel with
  // pending CTAs holds back every later kernel of the step (measured: 20 us holes in the decoder backward that runs next
  // to this pass) -- and a third of the SMs stays free for those kernels.
  const int nwave = std::max(1, ceil_div(ntile * rsplit, sm_count()));
  const int chunk = ceil_div(ntile, nwave);
  for (int x0 = 0; x0 < ntile && !rc; x0 += chunk) {
    p.x0 = x0;
    const dim3 grid(std::min(chunk, ntile - x0), rsplit);
    rc = a.zip ? launch(prior_bwd_tc_kernel<true, true>, grid, mm, mz, mzt)
               : launch(prior_bwd_tc_kernel<true, false>, grid, mm, mz, mzt);
  }
  if (rc || rsplit == 1) return rc;
  const size_t sh = sizeof(float) * (128 * (a.NG + 1) + 4 * a.NG);
  prior_bwd_cols_kernel<<<ntile, 512, sh, st>>>(a.gcol_part, a.tot_part, a.ms, a.isig, rsplit, a.C, a.D, a.LD, a.NG, a.Cpad,
                                               a.dmu, a.coldot_part);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? EXVAE_OK : (int)e;
}

}  // namespace exvae
