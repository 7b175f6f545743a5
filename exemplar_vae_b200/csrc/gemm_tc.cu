// K3 on the 5th-generation tensor cores: error-compensated 3xTF32 GEMM (sm_100a, tcgen05 + TMEM + TMA).
//
// The 1e-4 ELBO parity bar rules out single-pass TF32/BF16 operands, so every fp32
// operand element x is used as hi + lo with hi = x truncated to tf32 (what the tensor core does to the raw
// bits anyway, measured) and lo = tf32(x - hi), and each k-step issues three MMAs into the same TMEM
// accumulator:  D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi   (the dropped lo*lo term is ~2^-21).
// The split happens IN the kernel: the operands stay plain fp32 in HBM and every SM ingests 4 bytes per
// element (pre-split planes cost 8 and an extra HBM pass).
//
// Kernel anatomy (persistent: one CTA per SM walks 128 x BN output tiles; 448 threads):
//   warp 0       TMA producer: per 32-wide k-block, bulk-tensor loads of the fp32 A and B tiles into a
//                4-stage shared-memory ring that runs continuously across tiles
//   warps 2-9    converters: B: store lo at the same swizzled offset of the stage's lo slot (the raw
//                tile is the hi operand); A: thread = tile row, raw bits + lo -> tensor memory
//                (tcgen05.st); fences, then arrive on the stage's "converted" barrier
//   warp 1       allocates the 512 TMEM columns (two accumulators + 4 x 64 A columns); one elected lane
//                issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8; A from TMEM, B from shared memory)
//                x 4 k-steps x 3 products per stage; tcgen05.commit frees the stage / publishes the
//                finished accumulator
//   warps 10-13  epilogue of tile j while the MMA warp works on tile j+1: tcgen05.ld (32 lanes x 32
//                columns) -> bias / activation / gate -> shared-memory transposition patch -> 128-byte
//                row-segment stores; warp%4 = TMEM lane quadrant
// Both operand majors are supported straight from row-major global memory, so the three GEMMs of a
// dense layer need no transposed copies:
//   forward  x[R,K] W[O,K]^T              A K-major,  B K-major
//   dx       dY[R,O'] W[O',K]             A K-major,  B MN-major
//   dW       dY[R,O']^T x[R,K]            A MN-major, B MN-major   (split over R across CTAs)
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "tc_common.cuh"

namespace exvae {
using namespace tc;
namespace {

constexpr int TBM = 128;        // rows per CTA tile (UMMA M)
constexpr int TBK = 32;         // fp32 elements per k-block = one 128-byte swizzle row
constexpr int TSTAGES = 4;
constexpr int NCONV = 8;        // converter warps
constexpr int NEPI = 4;         // epilogue warps (one per TMEM lane quadrant)
constexpr int EPI_WARP0 = 2 + NCONV;
constexpr int TTHREADS = 32 * (2 + NCONV + NEPI);
constexpr int EP_LD = 36;       // row pitch (floats) of an epilogue warp's 32x32 transposition patch: 16-byte aligned rows,
                                // conflict-free for both the row-wise float4 writes and the 4-rows-per-instruction reads
constexpr int EP_BYTES = NEPI * 32 * EP_LD * 4;

struct TcParams {
  int M, N, K, kchunk;
  int ntm, ntn, ntiles;        // tile grid: ntiles = ntn * ntm * splits, n fastest
  int neff_last, goff_last;    // last column tile: UMMA N (16-multiple covering its valid columns) and, gated, the
                               // row offset of the g half inside the B tile (8-multiple covering the valid h columns)
  int gated_O;
  const float* bias0; const float* bias1;
  float* out0; float* out1; float* out2; int ldc;
  int act; float lo, hi;
  int c_vec;
  TcPriorEpi prior;            // TC_LSE / TC_PW epilogues
  // implicit-GEMM convolution (CONV): tile = bn images x bh output rows x OW pixels
  int cv_OH, cv_OW, cv_KW, cv_stride, cv_pad, cv_bh, cv_bn, cv_kbpt, cv_tpi, cv_mrows;
  int dwc_cin, dwc_kw, dwc_wp;   // implicit conv weight gradient (TcGemm): tap-shifted rows of the MN-major B operand
  unsigned long long* trace;   // debug: per-CTA {start, end, tiles, SM} (tools/gemm_trace.py), null in production
};
unsigned long long* g_trace = nullptr;
// CTA-pair (cta_group::2) variant of the big forward GEMMs: EXVAE_GEMM_PAIR=0 disables it
inline bool pair_enabled() {
  static const bool on = [] { const char* e = getenv("EXVAE_GEMM_PAIR"); return !(e && strcmp(e, "0") == 0); }();
  return on;
}

// debug (trace runs only): cycles spent inside a barrier wait
template <typename F>
__device__ __forceinline__ void timed_wait(bool on, long long& acc, F&& wait) {
  if (on) {
    const long long t0 = clock64();
    wait();
    acc += clock64() - t0;
  } else {
    wait();
  }
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// branch-free (expf is; an IEEE division is not: its slow-path call per element serialises the epilogue's 32
// independent columns): reciprocal to ~1 ulp, exact 0 for x < -88
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + expf(-x)); }

// The tensor core reads a tf32 operand by DROPPING the 13 low mantissa bits of the 32-bit container (measured on
// B200: a lo part taken relative to round-to-nearest gives 6e-4 errors, relative to truncation 5e-6), so the raw fp32
// tile in the hi slot already IS the hi operand: hi = x & ~0x1fff.  Only lo = tf32_rna(x - hi) has to be written
// (x - hi is exact; adding half an ulp of the 10-bit mantissa to the bit pattern and clearing the low bits is
// round-to-nearest, ties away).  A non-finite x stays non-finite in hi and so poisons the product.
__device__ __forceinline__ float split_lo(float x) {
  const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  const float r = x - h;
  return __uint_as_float((__float_as_uint(r) + 0x1000u) & 0xffffe000u);
}

// neff = UMMA N of this tile (BN except for a narrow last column tile: N = 300 is 2.34 tiles of 128), goff = first g row
// / TMEM column of the gated tile's second half, bbytes = bytes TMA delivers for the B tile, last = narrow tile maps
struct TileCoord {
  int m0, n0, kbeg, nkb, z, neff, goff, bbytes, last;
  int mrows;            // rows of the 128-row tile that hold data (implicit conv tiles may be shorter)
  int cn0, coh0;        // implicit conv: first image / first output row of the tile
};
template <int BN, int EPI, bool B_MN_, int CONV = 0, bool PAIR = false>
__device__ __forceinline__ TileCoord tile_coord(const TcParams& p, int t, int rank = 0) {
  TileCoord c;
  const int nt = t % p.ntn;
  const int rest = t / p.ntn;
  int mt = rest % p.ntm;
  c.z = rest / p.ntm;
  if (PAIR) mt = 2 * mt + rank;          // CTA pair: a 256-row tile = two 128-row units, one per CTA
  c.m0 = mt * TBM;
  c.mrows = TBM;
  c.cn0 = c.coh0 = 0;
  if (CONV) {
    if (p.cv_bn == 1) {
      c.cn0 = mt / p.cv_tpi;
      c.coh0 = (mt - c.cn0 * p.cv_tpi) * p.cv_bh;
    } else {
      c.cn0 = mt * p.cv_bn;
    }
    c.m0 = (c.cn0 * p.cv_OH + c.coh0) * p.cv_OW;
    c.mrows = p.cv_mrows;
  }
  c.n0 = (EPI == TC_GATED) ? nt * (BN / 2) : nt * BN;
  c.last = nt == p.ntn - 1;
  c.neff = c.last ? p.neff_last : BN;
  c.goff = c.last ? p.goff_last : BN / 2;
  c.bbytes = B_MN_ ? ((c.neff + 31) / 32) * 4096 : c.neff * 128;
  if (PAIR) c.bbytes >>= 1;              // each CTA of a pair holds half of the B tile's rows
  c.kbeg = c.z * p.kchunk;
  const int kend = min(p.K, c.kbeg + p.kchunk);
  c.nkb = max(0, (kend - c.kbeg + TBK - 1) / TBK);
  return c;
}

// Persistent: one CTA per SM walks the tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...; the shared-memory ring and
// its barriers run continuously across tiles (no pipeline drain), and two TMEM accumulators let the epilogue of tile
// j overlap the main loop of tile j+1.  Operands arrive as plain fp32 (TMA); the raw B tile is the hi operand and the
// converter warps store its lo part at the same (swizzled) offset of the stage's "lo" slot.
//
// The A operand goes through TENSOR MEMORY instead of shared memory.  The converter thread that owns TMEM lane m reads
// row m of the raw A tile (16 k-values: the two warps of a lane quadrant split the 32-wide k-block) and writes the hi
// part (the raw bits) and the lo part with tcgen05.st into the stage's 64 TMEM columns; the MMAs take A from TMEM.
// The A tile is then read from shared memory ONCE per k-block (instead of once by the converter + once per MMA, 3 x)
// and its lo part never touches shared memory: 128 KB instead of 192 KB of shared-memory traffic per k-block, which
// is what bounds this kernel (measured: TMA-only 0.33, + conversion 0.59, + MMAs 0.91 us per k-block, additive;
// profiles/r1_gemm_persistent.md).  A stage is 48 KB, so the ring has 4 of them.  An MN-major A tile needs no
// swizzle (no MMA reads it): one {128 m, 32 k} box, column m read by lane m without bank conflicts.
//
// PAIR: two CTAs of a cluster (one TPC) cooperate on a 256 x BN tile with tcgen05.mma.cta_group::2.  Each CTA loads and
// converts its own 128 rows of A (into its own TMEM) and HALF of the B tile's rows; the leader CTA issues M = 256 MMAs
// that read A from both tensor memories and B from both shared memories and write each CTA's 128 accumulator rows.
// The B operand is then staged, converted and read once per 256 output rows: 80 KB instead of 128 KB of shared-memory
// traffic per k-block and SM, below the 0.39 us the MMAs themselves need.
template <int BN, bool A_MN, bool B_MN, int EPI, int CONV = 0, bool PAIR = false>
__global__ void __launch_bounds__(TTHREADS, 1)
    gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmBl, const TcParams p) {
  static_assert(!PAIR || (!A_MN && !B_MN && (EPI == TC_GATED || EPI == TC_BIAS_ACT)), "pair mode: K-major forward GEMMs");
  constexpr int A_BYTES = TBM * TBK * 4;   // 16 KB: raw A tile
  constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * TBK * 4;    // raw (= hi) B tile (this CTA's rows); the lo tile follows it
  constexpr int STAGE_BYTES = A_BYTES + 2 * B_BYTES;
  constexpr int B_OFF = A_BYTES;                         // B hi slot inside a stage
  // Exemplar-prior epilogues keep TWO accumulators per tile: the hi*hi products in one, the two small cross products
  // (A_lo*B_hi + A_hi*B_lo, 2^-11 of the main term) in the other, added in fp32 by the epilogue.  The TMEM accumulate
  // truncates, so the error grows with (number of adds) x ulp(|accumulator|); logits are differences of terms ~|mu/sigma|^2
  // (1e3 at D=128), and keeping the cross terms out of the big accumulator cuts the truncating adds on it by 3.
  constexpr bool DUAL = (EPI == TC_LSE || EPI == TC_PW);
  constexpr int ACC_COLS = DUAL ? 2 * BN : BN;           // TMEM columns per accumulator buffer
  constexpr uint32_t TM_COLS = 512;                      // 2 accumulator buffers + TSTAGES x (32 hi + 32 lo) A columns
  constexpr uint32_t TM_A0 = 2 * ACC_COLS;
  static_assert(2 * ACC_COLS + TSTAGES * 64 <= 512, "TMEM budget");
  extern __shared__ unsigned char smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte aligned bases
  unsigned char* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TSTAGES * STAGE_BYTES);
  uint64_t* full = bars;                    // TMA bytes landed               (producer -> converters)
  uint64_t* conv = bars + TSTAGES;          // hi/lo planes written            (converters -> MMA)
  uint64_t* empty = bars + 2 * TSTAGES;     // MMAs that read the stage done   (MMA -> producer)
  uint64_t* acc_full = bars + 3 * TSTAGES;  // accumulator complete            (MMA -> epilogue)       [2]
  uint64_t* acc_empty = acc_full + 2;       // accumulator drained             (epilogue -> MMA)       [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = PAIR ? (int)cluster_ctarank() : 0;          // 0 = leader of the pair
  const int cta0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // first tile / tile stride of this CTA (pair)
  const int nct = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  unsigned long long* tr = p.trace ? p.trace + 8 * (size_t)blockIdx.x : nullptr;
  if (tr && tid == 0) tr[0] = gtimer();

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TSTAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&conv[s], PAIR ? 2 * NCONV : NCONV);      // pair: the leader's barrier collects both CTAs' converters
      mbar_init(&empty[s], 1);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], PAIR ? 2 * NEPI : NEPI);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc2(tmem_slot, TM_COLS);
    else tmem_alloc(tmem_slot, TM_COLS);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();          // the peer's barriers exist before anything arrives on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one_sync()) {
      int it = 0;
      long long w_prod = 0;
      for (int t = cta0; t < p.ntiles; t += nct) {
        const TileCoord c = tile_coord<BN, EPI, B_MN, CONV, PAIR>(p, t, rank);
        for (int kb = 0; kb < c.nkb; ++kb, ++it) {
          const int s = it % TSTAGES, ph = (it / TSTAGES) & 1;
          timed_wait(tr != nullptr, w_prod, [&] { mbar_wait(&empty[s], ph ^ 1); });
          mbar_arrive_expect_tx(&full[s], (CONV ? c.mrows * 128 : A_BYTES) + c.bbytes);
          unsigned char* sa = smem + s * STAGE_BYTES;
          unsigned char* sb = sa + B_OFF;
          const int k0 = c.kbeg + kb * TBK;
          if (CONV) {
            // k-block kb = (tap, 32-channel chunk): ONE box {32 c, OW pixels, bh rows, bn images} of the NHWC input,
            // shifted by the tap; TMA zero-fills the padding border and the channels beyond C
            const int tap = kb / p.cv_kbpt, cb = kb - tap * p.cv_kbpt;
            const int kh = tap / p.cv_KW, kw = tap - kh * p.cv_KW;
            tma_load_4d(sa, &tmA, &full[s], cb * 32, kw - p.cv_pad, c.coh0 * p.cv_stride + kh - p.cv_pad, c.cn0);
          } else if (!A_MN) {
            tma_load_2d(sa, &tmA, &full[s], k0, c.m0);                        // box {32 k, 128 rows}
          } else {
            tma_load_2d(sa, &tmA, &full[s], c.m0, k0);                        // box {128 m, 32 k}, no swizzle
          }
          if (PAIR) {
            // this CTA's half of the B tile's rows: gated = the h rows (leader) / the g rows (peer), else N/2 rows each
            const CUtensorMap* mb = c.last ? &tmBl : &tmB;
            if (EPI == TC_GATED) tma_load_2d(sb, mb, &full[s], k0, rank == 0 ? c.n0 : p.gated_O + c.n0);
            else tma_load_2d(sb, mb, &full[s], k0, c.n0 + rank * (c.neff >> 1));
          } else if (!B_MN) {
            const CUtensorMap* mb = c.last ? &tmBl : &tmB;                     // narrow last tile: smaller boxes
            if (EPI == TC_GATED) {                                             // box {32 k, goff rows}: h rows, g rows
              tma_load_2d(sb, mb, &full[s], k0, c.n0);
              tma_load_2d(sb + c.goff * 128, mb, &full[s], k0, p.gated_O + c.n0);
            } else {
              tma_load_2d(sb, mb, &full[s], k0, c.n0);                        // box {32 k, neff rows}
            }
          } else {
            const int nch = (c.neff + 31) / 32;
#pragma unroll
            for (int q = 0; q < BN / 32; ++q) {                                // box {32 n, 32 k}
              if (q < nch) {
                int nn = c.n0 + 32 * q, kk = k0;
                if (p.dwc_cin > 0) {          // column = (tap, channel): rows shifted by the tap inside the padded frame
                  const int tap = nn / p.dwc_cin;
                  nn -= tap * p.dwc_cin;
                  const int ky = tap / p.dwc_kw;
                  kk += ky * p.dwc_wp + (tap - ky * p.dwc_kw);
                }
                tma_load_2d(sb + q * 4096, &tmB, &full[s], nn, kk);
              }
            }
          }
        }
      }
      if (tr) tr[1] = (unsigned long long)w_prod;      // cycles the producer waited for a free stage
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (pair: the leader CTA only)
    if (rank == 0 && elect_one_sync()) {
      constexpr int UM = PAIR ? 2 * TBM : TBM;
      constexpr uint32_t idesc_full = umma_idesc(UM, BN, false, B_MN);    // A in TMEM is [m lanes][k columns]
      int it = 0, j = 0;
      long long w_mma = 0, w_acc = 0;
      for (int t = cta0; t < p.ntiles; t += nct, ++j) {
        const TileCoord c = tile_coord<BN, EPI, B_MN, CONV, PAIR>(p, t, rank);
        const int acc = j & 1;
        const uint32_t idesc = c.last ? umma_idesc(UM, c.neff, false, B_MN) : idesc_full;
        timed_wait(tr != nullptr, w_acc, [&] {
          if (PAIR) mbar_wait_cluster(&acc_empty[acc], ((j >> 1) & 1) ^ 1);
          else mbar_wait(&acc_empty[acc], ((j >> 1) & 1) ^ 1);      // the epilogue has drained this accumulator
        });
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
        const uint32_t d_cross = DUAL ? d_tmem + BN : d_tmem;
        for (int kb = 0; kb < c.nkb; ++kb, ++it) {
          const int s = it % TSTAGES, ph = (it / TSTAGES) & 1;
          timed_wait(tr != nullptr, w_mma, [&] {
            if (PAIR) mbar_wait_cluster(&conv[s], ph);
            else mbar_wait(&conv[s], ph);
          });
          tc_fence_after();
          const uint32_t sb_hi = smem_u32(smem + s * STAGE_BYTES) + B_OFF;
          const uint32_t sb_lo = sb_hi + B_BYTES;
          const uint32_t ta_hi = tmem_base + TM_A0 + (uint32_t)(s * 64);
#pragma unroll
          for (int ks = 0; ks < TBK / 8; ++ks) {
            // B K-major : advance 8 tf32 = 32 bytes inside the swizzled 128-byte row; LBO unused, SBO = 8 rows (1024 B)
            // B MN-major: advance 8 k-rows = 1024 bytes; LBO = next 32-wide MN chunk (4096 B), SBO = next group of
            //             4 k-rows (512 B) of the 32-byte-atom swizzle
            const uint32_t boff = B_MN ? ks * 1024 : ks * 32;
            const uint64_t b_hi = umma_desc(sb_hi + boff, B_MN ? 4096 : 16, B_MN ? 512 : 1024, B_MN ? 1 : 2);
            const uint64_t b_lo = umma_desc(sb_lo + boff, B_MN ? 4096 : 16, B_MN ? 512 : 1024, B_MN ? 1 : 2);
            const uint32_t first = (kb > 0 || ks > 0) ? 1u : 0u;
            if (PAIR) {
              umma_tf32_ts2(d_tmem, ta_hi + 32 + 8 * ks, b_hi, idesc, first);
              umma_tf32_ts2(d_tmem, ta_hi + 8 * ks, b_lo, idesc, 1u);
              umma_tf32_ts2(d_tmem, ta_hi + 8 * ks, b_hi, idesc, 1u);
            } else {
              umma_tf32_ts(d_cross, ta_hi + 32 + 8 * ks, b_hi, idesc, first);          // A_lo x B_hi: small terms first
              umma_tf32_ts(d_cross, ta_hi + 8 * ks, b_lo, idesc, 1u);                  // A_hi x B_lo
              umma_tf32_ts(d_tmem, ta_hi + 8 * ks, b_hi, idesc, DUAL ? first : 1u);    // A_hi x B_hi
            }
          }
          if (PAIR) umma_commit2(&empty[s]);      // frees the stage in BOTH CTAs once these MMAs have read it
          else umma_commit(&empty[s]);   // frees the stage once these MMAs have read it
        }
        if (PAIR) {
          umma_commit2(&acc_full[acc]);             // (pair GEMMs always have k-blocks)
        } else {
          if (c.nkb > 0) umma_commit(&acc_full[acc]);   // accumulator complete
          else mbar_arrive(&acc_full[acc]);
        }
      }
      if (tr) { tr[2] = j; tr[3] = (unsigned long long)w_mma; tr[4] = (unsigned long long)w_acc; }
    }
  } else if (warp < EPI_WARP0) {
    // ---------------------------------------------------------------- converters (warps 2..9)
    const int ct = tid - 64;                               // 0 .. 32*NCONV-1
    constexpr int CT = 32 * NCONV;
    constexpr int B_IT = B_BYTES / 16 / CT;
    static_assert(B_BYTES % (16 * CT) == 0, "B tile chunks must divide over the converters");
    static_assert(NCONV == 8, "two converter warps per TMEM lane quadrant split the k-block");
    const int qd = warp & 3, kh = (warp - 2) >> 2;         // TMEM lane quadrant, half of the k-block
    const int am = 32 * qd + lane;                         // tile row (= TMEM lane) of this thread
    int it = 0;
    long long w_conv = 0;
    for (int t = cta0; t < p.ntiles; t += nct) {
      const TileCoord c = tile_coord<BN, EPI, B_MN, CONV, PAIR>(p, t, rank);
      for (int kb = 0; kb < c.nkb; ++kb, ++it) {
        const int s = it % TSTAGES, ph = (it / TSTAGES) & 1;
        timed_wait(tr != nullptr, w_conv, [&] { mbar_wait(&full[s], ph); });
        unsigned char* sa = smem + s * STAGE_BYTES;
        unsigned char* sb = sa + B_OFF;
        const int bchunks = c.bbytes >> 4;                   // 16-byte chunks of the B tile that hold data
        float4 vb[B_IT];
#pragma unroll
        for (int i = 0; i < B_IT; ++i)
          if (ct + CT * i < bchunks) vb[i] = *reinterpret_cast<const float4*>(sb + (ct + CT * i) * 16);
        {
          uint32_t hi[16], lo[16];
          if (!A_MN) {
            // SWIZZLE_128B: 16-byte chunk j of row m sits at chunk position j ^ (m & 7) of the row's 128 bytes
            const unsigned char* rowp = sa + am * 128;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 v = *reinterpret_cast<const float4*>(rowp + (((4 * kh + i) ^ (am & 7)) << 4));
              hi[4 * i] = __float_as_uint(v.x); hi[4 * i + 1] = __float_as_uint(v.y);
              hi[4 * i + 2] = __float_as_uint(v.z); hi[4 * i + 3] = __float_as_uint(v.w);
            }
          } else {
#pragma unroll
            for (int kk = 0; kk < 16; ++kk)
              hi[kk] = *reinterpret_cast<const uint32_t*>(sa + (16 * kh + kk) * (TBM * 4) + am * 4);
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) lo[e] = __float_as_uint(split_lo(__uint_as_float(hi[e])));
          const uint32_t ta = tmem_base + ((uint32_t)(32 * qd) << 16) + TM_A0 + (uint32_t)(s * 64 + 16 * kh);
          tmem_st16(ta, hi);
          tmem_st16(ta + 32, lo);
          tmem_st_wait();
          tc_fence_before();
        }
#pragma unroll
        for (int i = 0; i < B_IT; ++i)
          if (ct + CT * i < bchunks)
            *reinterpret_cast<float4*>(sb + B_BYTES + (ct + CT * i) * 16) =
                make_float4(split_lo(vb[i].x), split_lo(vb[i].y), split_lo(vb[i].z), split_lo(vb[i].w));
        fence_proxy_async();        // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(&conv[s], 0);      // on the leader's barrier
          else mbar_arrive(&conv[s]);
        }
      }
    }
    if (tr && tid == 64) tr[5] = (unsigned long long)w_conv;   // cycles a converter warp waited for TMA data
  } else {
    // ---------------------------------------------------------------- epilogue (warps 10..13)
    // TMEM gives every thread one output ROW (32 consecutive columns per tcgen05.ld); storing that directly costs
    // one 16-byte request per thread per instruction (measured: ~11 us per tile, the bottleneck).  Each warp
    // therefore transposes its 32x32 sub-tile through a private shared-memory patch and writes whole 128-byte row
    // segments: 4 rows x 128 B per store instruction.
    const int q = warp & 3;       // TMEM lane quadrant this warp may access
    float* patch = reinterpret_cast<float*>(smem + TSTAGES * STAGE_BYTES + 256) + (warp - EPI_WARP0) * (32 * EP_LD);
    int j = 0;
    int tile_m0 = 0, tile_rows = TBM;     // current tile: first row and number of rows that hold data
    // write the staged 32x32 patch to dst[(row0 + r) * ldc + col0 + c] for c < ncols
    auto flush = [&](float* dst, int row0, int col0, int ncols) {
      const int ldc = (EPI == TC_PW) ? p.prior.ldw : p.ldc;
      __syncwarp();
      if (p.c_vec) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = 4 * i + (lane >> 3), ch = (lane & 7) * 4;
          const float4 v = *reinterpret_cast<const float4*>(patch + r * EP_LD + ch);
          if (row0 + r < p.M && row0 - tile_m0 + r < tile_rows) {
            float* d = dst + (size_t)(row0 + r) * ldc + col0 + ch;
            if (ch + 3 < ncols) {
              *reinterpret_cast<float4*>(d) = v;
            } else {
              if (ch < ncols) d[0] = v.x;
              if (ch + 1 < ncols) d[1] = v.y;
              if (ch + 2 < ncols) d[2] = v.z;
            }
          }
        }
      } else {
        if (lane < ncols)
          for (int r = 0; r < 32 && row0 + r < p.M && row0 - tile_m0 + r < tile_rows; ++r)
            dst[(size_t)(row0 + r) * ldc + col0 + lane] = patch[r * EP_LD + lane];
      }
      __syncwarp();
    };
    for (int t = cta0; t < p.ntiles; t += nct, ++j) {
      const TileCoord c = tile_coord<BN, EPI, B_MN, CONV, PAIR>(p, t, rank);
      const int acc = j & 1;
      mbar_wait(&acc_full[acc], (j >> 1) & 1);
      tc_fence_after();
      tile_m0 = c.m0;
      tile_rows = c.mrows;
      const int row0 = c.m0 + 32 * q;
      const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * ACC_COLS);
      float* prow = patch + lane * EP_LD;
      if (EPI == TC_GATED) {
#pragma unroll 1
        for (int c0 = 0; c0 < BN / 2; c0 += 32) {
          const int col0 = c.n0 + c0;
          if (col0 >= p.gated_O) break;   // warp-uniform
          const int ncols = min(32, p.gated_O - col0);
          uint32_t hv[32], gv[32];
          if (c.nkb > 0) {
            tmem_ld32(lane_addr + c0, hv);
            tmem_ld32(lane_addr + c.goff + c0, gv);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) hv[e] = gv[e] = 0u;
          }
          // bias of column col0 + lane, broadcast by shuffle
          const float bh_l = (lane < ncols && p.bias0) ? p.bias0[col0 + lane] : 0.f;
          const float bg_l = (lane < ncols && p.bias1) ? p.bias1[col0 + lane] : 0.f;
          float sg[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float hh = __uint_as_float(hv[e]) + __shfl_sync(0xffffffffu, bh_l, e);
            sg[e] = sigmoidf_(__uint_as_float(gv[e]) + __shfl_sync(0xffffffffu, bg_l, e));
            hv[e] = __float_as_uint(hh);
          }
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            *reinterpret_cast<float4*>(prow + e) =
                make_float4(__uint_as_float(hv[e]) * sg[e], __uint_as_float(hv[e + 1]) * sg[e + 1],
                            __uint_as_float(hv[e + 2]) * sg[e + 2], __uint_as_float(hv[e + 3]) * sg[e + 3]);
          flush(p.out0, row0, col0, ncols);
          if (p.out2) {
#pragma unroll
            for (int e = 0; e < 32; e += 4)
              *reinterpret_cast<float4*>(prow + e) = make_float4(sg[e], sg[e + 1], sg[e + 2], sg[e + 3]);
            flush(p.out2, row0, col0, ncols);
          }
          if (p.out1) {
#pragma unroll
            for (int e = 0; e < 32; e += 4)
              *reinterpret_cast<float4*>(prow + e) = make_float4(__uint_as_float(hv[e]), __uint_as_float(hv[e + 1]),
                                                                 __uint_as_float(hv[e + 2]), __uint_as_float(hv[e + 3]));
            flush(p.out1, row0, col0, ncols);
          }
        }
      } else if (EPI == TC_LSE || EPI == TC_PW) {
        // exemplar-prior epilogues: the accumulator holds base-2 logits S[row, col]
        const int row = row0 + lane;
        const bool rowok = row < p.M;
        const bool mask = p.prior.cidx != nullptr;
        const long long zi = (mask && rowok) ? p.prior.zidx[row] : (long long)0x8000000000000000ull;
        const int zlo = (int)zi;
        float m = -INFINITY, ssum = 0.f, cnt = 0.f;
        float gi = 0.f, li = INFINITY;
        if (EPI == TC_PW && rowok) { gi = p.prior.g[row]; li = p.prior.lse2[row]; }
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          const int col0 = c.n0 + c0;
          if (col0 >= p.N) break;   // warp-uniform
          const int ncols = min(32, p.N - col0);
          uint32_t v[32];
          {
            uint32_t vc[32];
            tmem_ld32(lane_addr + c0, v);              // hi*hi accumulator
            tmem_ld32(lane_addr + BN + c0, vc);        // cross-term accumulator
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(vc[e]));
          }
          int clo = 0;
          if (mask) clo = lane < ncols ? (int)p.prior.cidx[col0 + lane] : 0;
          if (EPI == TC_LSE) {
            float mx = -INFINITY;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              float x = e < ncols ? __uint_as_float(v[e]) : -INFINITY;
              if (mask) {
                const int ce = __shfl_sync(0xffffffffu, clo, e);
                if (ce == zlo && e < ncols && rowok && p.prior.cidx[col0 + e] == zi) {   // 32-bit pre-test, rare hit
                  x = -INFINITY;
                  cnt += 1.f;
                }
              }
              v[e] = __float_as_uint(x);
              mx = fmaxf(mx, x);
            }
            const float mn = fmaxf(m, mx);
            const float msafe = (mn == -INFINITY) ? 0.f : mn;
            float acc = ssum * ex2_approx(m - msafe);
#pragma unroll
            for (int e = 0; e < 32; ++e) acc += ex2_approx(__uint_as_float(v[e]) - msafe);
            ssum = acc;
            m = mn;
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              float w = gi * ex2_approx(__uint_as_float(v[e]) - li);
              if (mask) {
                const int ce = __shfl_sync(0xffffffffu, clo, e);
                if (ce == zlo && e < ncols && rowok && p.prior.cidx[col0 + e] == zi) w = 0.f;
              }
              if (li == -INFINITY || e >= ncols) w = 0.f;
              v[e] = __float_as_uint(w);
              // transposed copy: lane = row, so one store instruction covers 32 consecutive rows of column col0+e
              if (rowok && e < ncols) p.prior.wt[(size_t)(col0 + e) * p.prior.ldwt + row] = w;
            }
#pragma unroll
            for (int e = 0; e < 32; e += 4)
              *reinterpret_cast<float4*>(prow + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                                 __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
            flush(p.prior.w, row0, col0, ncols);
          }
        }
        if (EPI == TC_LSE && rowok)
          reinterpret_cast<float4*>(p.prior.part)[(size_t)row * p.ntn + (c.n0 / BN)] = make_float4(m, ssum, cnt, 0.f);
      } else {
        float* dst_base = (EPI == TC_SPLITK) ? p.out0 + (size_t)c.z * p.M * p.ldc : p.out0;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          const int col0 = c.n0 + c0;
          if (col0 >= p.N) break;   // warp-uniform
          const int ncols = min(32, p.N - col0);
          uint32_t v[32];
          if (c.nkb > 0) {
            tmem_ld32(lane_addr + c0, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = 0u;
          }
          if (EPI == TC_BIAS_ACT) {
            const float b_l = (lane < ncols && p.bias0) ? p.bias0[col0 + lane] : 0.f;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              float x = __uint_as_float(v[e]) + __shfl_sync(0xffffffffu, b_l, e);
              if (p.act == EXVAE_ACT_SIGMOID) x = sigmoidf_(x);
              else if (p.act == EXVAE_ACT_HARDTANH) x = fminf(fmaxf(x, p.lo), p.hi);
              else if (p.act == EXVAE_ACT_RELU) x = fmaxf(x, 0.f);
              v[e] = __float_as_uint(x);
            }
          }
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            *reinterpret_cast<float4*>(prow + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                               __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
          flush(dst_base, row0, col0, ncols);
        }
      }
      // all tcgen05.ld of this accumulator have completed (wait::ld): hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(&acc_empty[acc], 0);
        else mbar_arrive(&acc_empty[acc]);
      }
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  if (tr && tid == 0) {
    unsigned int sm;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    tr[6] = gtimer();
    tr[7] = sm;
  }
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, TM_COLS);
    else tmem_dealloc(tmem_base, TM_COLS);
  }
}

// out[r, 0:K] = W0[r] for r < O, W1[r - O] for O <= r < 2*O: the gated layer's two weight tensors as one operand
__global__ void __launch_bounds__(256) concat2_kernel(const float* __restrict__ w0, const float* __restrict__ w1, size_t n4,
                                                      float* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < 2 * n4; i += (size_t)gridDim.x * blockDim.x)
    reinterpret_cast<float4*>(out)[i] = i < n4 ? reinterpret_cast<const float4*>(w0)[i] : reinterpret_cast<const float4*>(w1)[i - n4];
}

template <int BN, bool A_MN, bool B_MN, int EPI, int CONV = 0, bool PAIR = false>
int launch(const TcGemm& g, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mbl, TcParams& p,
           cudaStream_t st) {
  constexpr int STAGE = TBM * TBK * 4 + 2 * (PAIR ? BN / 2 : BN) * TBK * 4;
  constexpr int SMEM = TSTAGES * STAGE + 1024 + 256 + EP_BYTES;
  auto kern = gemm_tf32x3_kernel<BN, A_MN, B_MN, EPI, CONV, PAIR>;
  static bool configured = false;
  if (!configured) {
    EXVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  p.ntn = EPI == TC_GATED ? ceil_div(g.gated_O, BN / 2) : ceil_div(g.N, BN);
  p.ntm = ceil_div(g.M, TBM);
  if (CONV) p.ntm = g.conv->bn == 1 ? g.conv->N * (g.conv->OH / g.conv->bh) : ceil_div(g.conv->N, g.conv->bn);
  if (PAIR) p.ntm = ceil_div(p.ntm, 2);                  // 256-row tiles: two 128-row units per CTA pair
  p.ntiles = p.ntn * p.ntm * (EPI == TC_SPLITK ? g.splits : 1);
  if (PAIR) {
    const int pairs = std::min(p.ntiles, sm_count() / 2);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(TTHREADS);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ma, mb, mbl, p);
    return e == cudaSuccess ? EXVAE_OK : (int)e;
  }
  const int grid = std::min(p.ntiles, sm_count());     // persistent: one CTA per SM
  kern<<<grid, TTHREADS, SMEM, st>>>(ma, mb, mbl, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? EXVAE_OK : (int)e;
}

}  // namespace

bool tc_enabled() {
  static int state = -1;
  if (state < 0) {
    const char* env = getenv("EXVAE_GEMM");
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

bool tc_dims_ok(int pitch) { return pitch > 0 && pitch % 4 == 0; }

void tc_set_trace(unsigned long long* buf) { g_trace = buf; }
unsigned long long* tc_take_trace(size_t words) {
  unsigned long long* t = g_trace;
  if (g_trace) g_trace += words;
  return t;
}

int tc_concat2(const float* w0, const float* w1, size_t n, float* out, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(w0) & 15) || (reinterpret_cast<uintptr_t>(w1) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 15) || (n & 3))
    return EXVAE_ERR_INVALID_ARG;
  const size_t n4 = n / 4;
  const int blocks = (int)std::min<size_t>((2 * n4 + 255) / 256, 148 * 8);
  concat2_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, st>>>(w0, w1, n4, out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? EXVAE_OK : (int)e;
}

template <int BN>
static int tc_gemm_launch_bn(const TcGemm& g, cudaStream_t st) {
  CUtensorMap ma, mb, mbl;
  // A never feeds an MMA from shared memory: an MN-major tile is one un-swizzled {128 m, 32 k} box
  int rc;
  if (g.conv) {
    const TcConv& cv = *g.conv;
    rc = make_map_nhwc(&ma, cv.x, cv.N, cv.H, cv.W, cv.C, cv.OW, cv.bh, cv.bn, cv.stride);
  } else {
    rc = g.a_mn ? make_map2d(&ma, g.a, g.a_rows, g.a_cols, TBM, 32, CU_TENSOR_MAP_SWIZZLE_NONE)
                : make_map2d(&ma, g.a, g.a_rows, g.a_cols, TBM, false);
  }
  if (rc) return rc;
  rc = make_map2d(&mb, g.b, g.b_rows, g.b_cols, g.b_mn ? 32 : (g.epi == TC_GATED ? BN / 2 : BN), g.b_mn);
  if (rc) return rc;
  TcParams p{};
  // narrow last column tile: UMMA N = the 16-multiple that covers its valid columns (gated: both halves, 8 each)
  if (g.epi == TC_GATED) {
    const int h_last = g.gated_O - (ceil_div(g.gated_O, BN / 2) - 1) * (BN / 2);
    p.goff_last = ceil_div(h_last, 8) * 8;
    p.neff_last = 2 * p.goff_last;
  } else {
    const int n_last = g.N - (ceil_div(g.N, BN) - 1) * BN;
    p.neff_last = ceil_div(n_last, 16) * 16;
    p.goff_last = 0;
  }
  // K-major B: the last tile is loaded with boxes of exactly that many rows (MN-major B just loads fewer 32-wide chunks)
  rc = make_map2d(&mbl, g.b, g.b_rows, g.b_cols, g.b_mn ? 32 : (g.epi == TC_GATED ? p.goff_last : p.neff_last), g.b_mn);
  if (rc) return rc;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.kchunk = g.epi == TC_SPLITK ? g.kchunk : g.K;
  p.gated_O = g.gated_O;
  p.bias0 = g.bias0; p.bias1 = g.bias1; p.out0 = g.out0; p.out1 = g.out1; p.out2 = g.out2; p.ldc = g.ldc;
  p.act = g.act; p.lo = g.lo; p.hi = g.hi;
  p.prior = g.prior;
  if (g.conv) {
    const TcConv& cv = *g.conv;
    p.cv_OH = cv.OH; p.cv_OW = cv.OW; p.cv_KW = cv.KW; p.cv_stride = cv.stride; p.cv_pad = cv.pad;
    p.cv_bh = cv.bh; p.cv_bn = cv.bn; p.cv_kbpt = ceil_div(cv.C, 32); p.cv_tpi = cv.OH / cv.bh;
    p.cv_mrows = cv.bn * cv.bh * cv.OW;
  }
  p.dwc_cin = g.dwc_cin; p.dwc_kw = g.dwc_kw; p.dwc_wp = g.dwc_wp;
  p.trace = g_trace;
  if (g_trace) g_trace += 8 * 160;   // the next traced launch writes the next segment
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  p.c_vec = (g.ldc % 4 == 0) && al16(g.out0) && (!g.out1 || al16(g.out1)) && (!g.out2 || al16(g.out2)) &&
            (g.epi != TC_SPLITK || ((size_t)g.M * g.ldc) % 4 == 0);
  if (g.epi == TC_PW) p.c_vec = (g.prior.ldw % 4 == 0) && al16(g.prior.w);
  if constexpr (BN == 64) {      // the dual-accumulator prior epilogues only exist for 64-wide tiles (TMEM budget)
    if (g.epi == TC_LSE && !g.a_mn && !g.b_mn) return launch<BN, false, false, TC_LSE>(g, ma, mb, mbl, p, st);
    if (g.epi == TC_PW && !g.a_mn && !g.b_mn) return launch<BN, false, false, TC_PW>(g, ma, mb, mbl, p, st);
  }
  if constexpr (BN == 128) {
    // CTA pairs (cta_group::2) for the big K-major forward GEMMs: enough 256-row tiles to fill every pair of SMs
    if (pair_enabled() && !g.a_mn && !g.b_mn && (g.epi == TC_GATED || g.epi == TC_BIAS_ACT) && g.K >= 64) {
      const int ntn = g.epi == TC_GATED ? ceil_div(g.gated_O, BN / 2) : ceil_div(g.N, BN);
      const int units = g.conv ? (g.conv->bn == 1 ? g.conv->N * (g.conv->OH / g.conv->bh) : ceil_div(g.conv->N, g.conv->bn))
                               : ceil_div(g.M, TBM);
      if ((long long)ntn * ceil_div(units, 2) >= sm_count() / 2) {
        CUtensorMap mbp = mb, mblp = mbl;
        if (g.epi == TC_BIAS_ACT) {          // N/2 rows of the B tile per CTA
          rc = make_map2d(&mbp, g.b, g.b_rows, g.b_cols, BN / 2, false);
          if (rc) return rc;
          rc = make_map2d(&mblp, g.b, g.b_rows, g.b_cols, p.neff_last / 2, false);
          if (rc) return rc;
        }
        if (g.conv) {
          if (g.epi == TC_GATED) return launch<BN, false, false, TC_GATED, 1, true>(g, ma, mbp, mblp, p, st);
          return launch<BN, false, false, TC_BIAS_ACT, 1, true>(g, ma, mbp, mblp, p, st);
        }
        if (g.epi == TC_GATED) return launch<BN, false, false, TC_GATED, 0, true>(g, ma, mbp, mblp, p, st);
        return launch<BN, false, false, TC_BIAS_ACT, 0, true>(g, ma, mbp, mblp, p, st);
      }
    }
  }
  if (g.conv) {      // implicit-GEMM convolution: forward (gated / bias+activation) and the stride-1 input gradient (plain)
    if (g.a_mn || g.b_mn) return EXVAE_ERR_UNSUPPORTED;
    if (g.epi == TC_GATED) return launch<BN, false, false, TC_GATED, 1>(g, ma, mb, mbl, p, st);
    if (g.epi == TC_BIAS_ACT) return launch<BN, false, false, TC_BIAS_ACT, 1>(g, ma, mb, mbl, p, st);
    return EXVAE_ERR_UNSUPPORTED;
  }
  if (g.epi == TC_GATED && !g.a_mn && !g.b_mn) return launch<BN, false, false, TC_GATED>(g, ma, mb, mbl, p, st);
  if (g.epi == TC_BIAS_ACT && !g.a_mn && !g.b_mn) return launch<BN, false, false, TC_BIAS_ACT>(g, ma, mb, mbl, p, st);
  if (g.epi == TC_PLAIN && !g.a_mn && g.b_mn) return launch<BN, false, true, TC_PLAIN>(g, ma, mb, mbl, p, st);
  if (g.epi == TC_SPLITK && g.a_mn && g.b_mn) return launch<BN, true, true, TC_SPLITK>(g, ma, mb, mbl, p, st);
  return EXVAE_ERR_UNSUPPORTED;
}

static bool tc_use_bn64(const TcGemm& g) {
  if (g.epi == TC_LSE || g.epi == TC_PW) return true;     // two accumulators per tile: 2 x 2 x 64 TMEM columns
  if (g.epi == TC_SPLITK) return false;
  const int ntn = g.epi == TC_GATED ? ceil_div(g.gated_O, 64) : ceil_div(g.N, 128);
  return 2 * ceil_div(g.M, TBM) * ntn <= sm_count();
}
bool tc_conv_tiling(int OH, int OW, int* bh, int* bn) {
  if (OW > TBM || OW < 1 || OH < 1) return false;
  int best = 1;
  for (int d = 1; d <= OH; ++d)
    if (OH % d == 0 && d * OW <= TBM) best = d;
  *bh = best;
  *bn = best == OH ? std::max(1, TBM / (OH * OW)) : 1;
  if (*bn > 256) *bn = 256;
  return true;
}

int tc_gemm_ntn(const TcGemm& g) { return ceil_div(g.N, tc_use_bn64(g) ? 64 : 128); }

int tc_gemm_launch(const TcGemm& g, cudaStream_t st) {
  // Small GEMMs (the decoder and the batch-only heads: a few 128-wide tiles) are latency-bound and leave most SMs idle:
  // 64-wide tiles double the number of CTAs and halve each CTA's B-side work and epilogue.
  if (tc_use_bn64(g)) return tc_gemm_launch_bn<64>(g, st);
  return tc_gemm_launch_bn<128>(g, st);
}

}  // namespace exvae
