// K1 forward on the tensor cores (sm_100a): logit2 tiles by tcgen05.mma (3xTF32), online log-sum-exp
// straight out of TMEM.
//
// The staged operands are augmented so that ONE contraction yields the base-2 logit:
//     z'_b = (zs_b * log2e | 1 | 0..)   m'_n = (ms_n | -0.5*||ms_n||^2*log2e | 0..)   =>  z'_b . m'_n = logit2[b,n]
// (K = D+1 padded to a multiple of 8; D = 40 -> 48 = 6 k-steps of the tf32 UMMA_K = 8), each split
// into hi/lo tf32 planes: 3 MMAs per k-step keep the result inside the 1e-4 parity bar.
//
// CTA = one block of 128 latents x a range of 128-column bank tiles, 320 threads, 1 CTA/SM:
//   warp 0     TMA: the z' tile once (resident), bank tiles through a 2-stage ring
//   warp 1     MMA issuer: 18 tcgen05.mma per tile into one of two 128-column TMEM accumulators
//   warps 2-9  epilogue: thread = (row, 64-column half); tcgen05.ld -> release the accumulator ->
//              max / ex2 / sum in registers (4 instructions per pair), leave-one-out patches only in
//              tiles that contain one of the row's masked columns (list built by a pre-pass)
// MMA (576 clk/tile) and epilogue (~600 issue clk/tile/SMSP) overlap, so the kernel runs at the
// MUFU/ALU rate of the soft-max instead of the FMA rate of the contraction.
#include "prior_lse_tc.cuh"

#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "tc_common.cuh"

namespace exvae {
using namespace tc;
namespace {

constexpr int PT_BN = 128;
constexpr int PT_THREADS = 320;
constexpr int PT_EPI_WARPS = 8;
constexpr int PT_TILE_BYTES = 128 * 128;            // one [128 rows x 32 floats] swizzled box
constexpr int PT_MAXM = 8;            // must equal PR_MAXM in prior_lse.cu
constexpr long long kPad = INT64_MIN;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <bool MASK>
__global__ void __launch_bounds__(PT_THREADS, 1)
    prior_lse_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmM,
                            const int* __restrict__ mcnt, const int* __restrict__ mlist,
                            const int64_t* __restrict__ cidx, const int64_t* __restrict__ z_idx, int B, int C, int KP,
                            int ntile, int nsplit, float* __restrict__ part) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  // layout: A [2 kb][2 planes] 64 KB | B stage 0 64 KB | B stage 1 64 KB | red | barriers
  unsigned char* sA = smem;
  unsigned char* sB = smem + 4 * PT_TILE_BYTES;
  float2* red = reinterpret_cast<float2*>(smem + 12 * PT_TILE_BYTES);          // [128] (m, s) of column half 1
  // column-index ring, 4 deep: the producer runs at most 2 tiles ahead of the MMA and the MMA at most 2
  // ahead of the epilogue (which releases its accumulator only after using the indices), so slot it%4 is
  // never overwritten while still being read and needs no barrier of its own.
  long long* cis = reinterpret_cast<long long*>(smem + 12 * PT_TILE_BYTES + 1024);   // [4][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 12 * PT_TILE_BYTES + 1024 + 4096);
  uint64_t* a_full = bars;
  uint64_t* b_full = bars + 1;      // [2]
  uint64_t* b_empty = bars + 3;     // [2]
  uint64_t* acc_full = bars + 5;    // [2]
  uint64_t* acc_empty = bars + 7;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x, rb = blockIdx.y;
  const int t0 = (int)(((long long)ntile * split) / nsplit);
  const int t1 = (int)(((long long)ntile * (split + 1)) / nsplit);
  const int nkb = (KP + 31) / 32;
  const int nks = KP / 8;

  if (tid == 0) {
    mbar_init(a_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], PT_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(a_full, nkb * 2 * PT_TILE_BYTES);
      for (int kb = 0; kb < nkb; ++kb)
        for (int pl = 0; pl < 2; ++pl) tma_load_3d(sA + (kb * 2 + pl) * PT_TILE_BYTES, &tmZ, a_full, kb * 32, rb * 128, pl);
      for (int t = t0; t < t1; ++t) {
        const int it = t - t0, s = it & 1, ph = (it >> 1) & 1;
        mbar_wait(&b_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&b_full[s], nkb * 2 * PT_TILE_BYTES + (MASK ? PT_BN * 8 : 0));
        for (int kb = 0; kb < nkb; ++kb)
          for (int pl = 0; pl < 2; ++pl)
            tma_load_3d(sB + (s * 4 + kb * 2 + pl) * PT_TILE_BYTES, &tmM, &b_full[s], kb * 32, t * PT_BN, pl);
        if (MASK) bulk_g2s(cis + (it & 3) * PT_BN, cidx + (size_t)t * PT_BN, PT_BN * 8, &b_full[s]);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc(128, PT_BN, false, false);
      mbar_wait(a_full, 0);
      for (int t = t0; t < t1; ++t) {
        const int it = t - t0, s = it & 1, ph = (it >> 1) & 1;
        mbar_wait(&b_full[s], ph);
        mbar_wait(&acc_empty[s], ph ^ 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB + s * 4 * PT_TILE_BYTES);
        for (int ks = 0; ks < nks; ++ks) {
          const int kb = ks >> 2, kk = ks & 3;
          const uint32_t off_hi = (kb * 2 + 0) * PT_TILE_BYTES + kk * 32, off_lo = (kb * 2 + 1) * PT_TILE_BYTES + kk * 32;
          const uint64_t a_hi = umma_desc(a0 + off_hi, 16, 1024, 2), a_lo = umma_desc(a0 + off_lo, 16, 1024, 2);
          const uint64_t b_hi = umma_desc(b0 + off_hi, 16, 1024, 2), b_lo = umma_desc(b0 + off_lo, 16, 1024, 2);
          const uint32_t d = tmem_base + s * PT_BN;
          umma_tf32(d, a_lo, b_hi, idesc, ks > 0 ? 1u : 0u);
          umma_tf32(d, a_hi, b_lo, idesc, 1u);
          umma_tf32(d, a_hi, b_hi, idesc, 1u);
        }
        umma_commit(&b_empty[s]);
        umma_commit(&acc_full[s]);
      }
    }
  } else {
    // ------------------------------------------------------------- epilogue warps 2..9
    const int q = warp & 3;                    // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;          // which 64 of the tile's 128 columns
    const int r = 32 * q + lane;               // row inside the block
    const int b = rb * 128 + r;
    int cnt = 0, ml[PT_MAXM];
#pragma unroll
    for (int k = 0; k < PT_MAXM; ++k) ml[k] = -1;
    long long zi = kPad;
    if (MASK && b < B) {
      cnt = mcnt[b];
#pragma unroll
      for (int k = 0; k < PT_MAXM; ++k)
        if (k < cnt) ml[k] = mlist[(size_t)b * PT_MAXM + k];
      zi = z_idx[b];
    }
    float m = -INFINITY, ssum = 0.f;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * q) << 16) + half * 64;
    for (int t = t0; t < t1; ++t) {
      const int it = t - t0, s = it & 1, ph = (it >> 1) & 1;
      mbar_wait(&acc_full[s], ph);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(lane_addr + s * PT_BN, v0);
      tmem_ld32(lane_addr + s * PT_BN + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      const int c0 = t * PT_BN + half * 64;
      if (MASK && cnt > 0) {
        if (cnt <= PT_MAXM) {                             // the row's masked columns are all in the list
#pragma unroll
          for (int k = 0; k < PT_MAXM; ++k) {
            const int jj = ml[k] - c0;                    // unused slots hold -1
            if (jj >= 0 && jj < 64) {                     // rare: this tile half contains a masked pair
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (j == jj) v0[j] = 0xff800000u;
                if (j + 32 == jj) v1[j] = 0xff800000u;
              }
            }
          }
        } else {                                          // list overflow (pathological duplication): compare all
#pragma unroll 4
          for (int j = 0; j < 64; ++j) {
            const int col = c0 + j;
            if (col < C && cis[(it & 3) * PT_BN + half * 64 + j] == zi) {
#pragma unroll
              for (int jx = 0; jx < 32; ++jx) {
                if (jx == j) v0[jx] = 0xff800000u;
                if (jx + 32 == j) v1[jx] = 0xff800000u;
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[s]);        // accumulator (and index stage) drained
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
    // merge the two column halves of each row and emit the (max, sum, masked-count) partial
    if (half == 1) red[r] = make_float2(m, ssum);
    named_bar_sync(1, PT_EPI_WARPS * 32);
    if (half == 0) {
      const float2 o = red[r];
      lse2_merge(m, ssum, o.x, o.y);
      reinterpret_cast<float4*>(part)[(size_t)b * nsplit + split] =
          make_float4(m, ssum, (MASK && split == 0) ? (float)cnt : 0.f, 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

bool prior_tc_enabled() {
  static int state = -1;
  if (state < 0) {
    const char* env = getenv("EXVAE_PRIOR");
    bool on = !(env && strcmp(env, "simt") == 0);
    int dev = 0, major = 0;
    if (on && (cudaGetDevice(&dev) != cudaSuccess ||
               cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10))
      on = false;
    if (on && !encode_fn()) on = false;
    (void)cudaGetLastError();
    state = on ? 1 : 0;
  }
  return state == 1;
}

int prior_fwd_tc_launch(const PriorTcArgs& a, int* nsplit_out, cudaStream_t st) {
  CUtensorMap mz, mm;
  int rc = make_map(&mz, a.zp, a.Bpad, a.KP, 128, false);
  if (rc) return rc;
  rc = make_map(&mm, a.mp, a.Cpad, a.KP, 128, false);
  if (rc) return rc;
  const int ntile = ceil_div(a.C, PT_BN);
  const int rbs = a.Bpad / 128;
  int nsplit = std::max(1, sm_count() / rbs);
  nsplit = std::min(nsplit, ntile);
  *nsplit_out = nsplit;
  constexpr int SMEM = 12 * PT_TILE_BYTES + 1024 + 4096 + 128 + 1024;
  dim3 grid(nsplit, rbs);
  auto launch = [&](auto kern) -> int {
    EXVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    kern<<<grid, PT_THREADS, SMEM, st>>>(mz, mm, a.mcnt, a.mlist, a.cidx, a.z_idx, a.B, a.C, a.KP, ntile, nsplit,
                                         a.part);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? EXVAE_OK : (int)e;
  };
  return a.mcnt ? launch(prior_lse_fwd_tc_kernel<true>) : launch(prior_lse_fwd_tc_kernel<false>);
}

}  // namespace exvae
