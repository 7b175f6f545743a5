// K3 on the 5th-generation tensor cores: error-compensated 3xTF32 GEMM (sm_100a, tcgen05 + TMEM + TMA).
//
// The 1e-4 ELBO parity bar rules out single-pass TF32/BF16 operands, so every fp32
// operand is split once into hi = tf32(x) and lo = tf32(x - hi) (both exactly representable, so
// the tensor core's own operand truncation is irrelevant) and each k-step issues three MMAs into
// the same TMEM accumulator:  D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo   (the lo*lo term is 2^-22).
//
// Kernel anatomy (one 128 x BN output tile per CTA, 192 threads, 1 CTA/SM):
//   warp 0      TMA producer: per 32-wide k-block, bulk-tensor loads of the A_hi/A_lo/B_hi/B_lo
//               tiles (SWIZZLE_128B boxes) into a 3-stage shared-memory ring, mbarrier expect_tx
//   warp 1      allocates BN TMEM columns; one elected lane issues tcgen05.mma.kind::tf32
//               (M=128, N=BN, K=8) x 4 k-steps x 3 products per stage, tcgen05.commit frees the stage
//   warps 2-5   epilogue: tcgen05.ld (32 lanes x 32 columns per instruction) -> bias / activation /
//               gate -> global stores; the thread owning TMEM lane r owns output row r
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
constexpr int TSTAGES = 3;
constexpr int TTHREADS = 192;

struct TcParams {
  int M, N, K, kchunk;
  int gated_O;
  const float* bias0; const float* bias1;
  float* out0; float* out1; float* out2; int ldc;
  int act; float lo, hi;
  int c_vec;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// CL = thread-block cluster size along the M tiles: the CL CTAs of a cluster compute different row blocks of
// the SAME column tile, so each loads only 1/CL of the B tile and TMA-multicasts it to its peers.
template <int BN, bool A_MN, bool B_MN, int EPI, int CL>
__global__ void __launch_bounds__(TTHREADS, 1)
    gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  constexpr int A_BYTES = TBM * TBK * 4;   // 16 KB per plane
  constexpr int B_BYTES = BN * TBK * 4;
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  extern __shared__ unsigned char smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte aligned bases
  unsigned char* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TSTAGES * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + TSTAGES;
  uint64_t* tmem_full = bars + 2 * TSTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TSTAGES + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * TBM;
  const int n0 = (EPI == TC_GATED) ? blockIdx.x * (BN / 2) : blockIdx.x * BN;
  const int kbeg = blockIdx.z * p.kchunk;
  const int kend = min(p.K, kbeg + p.kchunk);
  const int nkb = (kend - kbeg + TBK - 1) / TBK;

  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
  constexpr uint16_t cmask = (uint16_t)((1u << CL) - 1);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TSTAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CL);      // every CTA of the cluster must have consumed the stage before it is refilled
    }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();    // peers' barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TSTAGES, ph = (kb / TSTAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
        unsigned char* st = smem + s * STAGE_BYTES;
        const int k0 = kbeg + kb * TBK;
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
          unsigned char* sa = st + pl * A_BYTES;
          if (!A_MN) {
            tma_load_3d(sa, &tmA, &full[s], k0, m0, pl);                    // box {32 k, 128 rows}
          } else {
#pragma unroll
            for (int c = 0; c < TBM / 32; ++c)                              // box {32 m, 32 k}
              tma_load_3d(sa + c * 4096, &tmA, &full[s], m0 + 32 * c, k0, pl);
          }
          unsigned char* sb = st + 2 * A_BYTES + pl * B_BYTES;
          if (CL == 1) {
            if (!B_MN) {
              if (EPI == TC_GATED) {                                           // box {32 k, BN/2 rows}
                tma_load_3d(sb, &tmB, &full[s], k0, n0, pl);
                tma_load_3d(sb + (BN / 2) * 128, &tmB, &full[s], k0, p.gated_O + n0, pl);
              } else {
                tma_load_3d(sb, &tmB, &full[s], k0, n0, pl);                  // box {32 k, BN rows}
              }
            } else {
#pragma unroll
              for (int c = 0; c < BN / 32; ++c)                                // box {32 n, 32 k}
                tma_load_3d(sb + c * 4096, &tmB, &full[s], n0 + 32 * c, k0, pl);
            }
          } else {
            // this CTA's 1/CL share of the B tile, multicast to the whole cluster
            constexpr int SHARE = BN / CL;                                     // rows (K-major) or columns (MN-major)
            const int r0 = crank * SHARE;                                      // first tile row/column of the share
            if (!B_MN) {                                                       // box {32 k, SHARE rows}
              int grow = n0 + r0;
              if (EPI == TC_GATED) grow = (r0 < BN / 2) ? n0 + r0 : p.gated_O + n0 + (r0 - BN / 2);
              tma_load_3d_mc(sb + r0 * 128, &tmB, &full[s], k0, grow, pl, cmask);
            } else {
#pragma unroll
              for (int c = 0; c < SHARE / 32; ++c)                             // box {32 n, 32 k}
                tma_load_3d_mc(sb + (r0 / 32 + c) * 4096, &tmB, &full[s], n0 + r0 + 32 * c, k0, pl, cmask);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(TBM, BN, A_MN, B_MN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TSTAGES, ph = (kb / TSTAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t sa_hi = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t sa_lo = sa_hi + A_BYTES;
        const uint32_t sb_hi = sa_hi + 2 * A_BYTES;
        const uint32_t sb_lo = sb_hi + B_BYTES;
#pragma unroll
        for (int ks = 0; ks < TBK / 8; ++ks) {
          // K-major : advance 8 tf32 = 32 bytes inside the swizzled 128-byte row; LBO unused, SBO = 8 rows (1024 B)
          // MN-major: advance 8 k-rows = 1024 bytes; LBO = next 32-wide MN chunk (4096 B), SBO = next group of
          //           4 k-rows (512 B) of the 32-byte-atom swizzle
          const uint32_t aoff = A_MN ? ks * 1024 : ks * 32;
          const uint32_t boff = B_MN ? ks * 1024 : ks * 32;
          const uint64_t a_hi = umma_desc(sa_hi + aoff, A_MN ? 4096 : 16, A_MN ? 512 : 1024, A_MN ? 1 : 2);
          const uint64_t a_lo = umma_desc(sa_lo + aoff, A_MN ? 4096 : 16, A_MN ? 512 : 1024, A_MN ? 1 : 2);
          const uint64_t b_hi = umma_desc(sb_hi + boff, B_MN ? 4096 : 16, B_MN ? 512 : 1024, B_MN ? 1 : 2);
          const uint64_t b_lo = umma_desc(sb_lo + boff, B_MN ? 4096 : 16, B_MN ? 512 : 1024, B_MN ? 1 : 2);
          umma_tf32(tmem_base, a_lo, b_hi, idesc, (kb > 0 || ks > 0) ? 1u : 0u);   // small terms first
          umma_tf32(tmem_base, a_hi, b_lo, idesc, 1u);
          umma_tf32(tmem_base, a_hi, b_hi, idesc, 1u);
        }
        // frees the stage (in every CTA of the cluster) once these MMAs have read it
        if (CL == 1) umma_commit(&empty[s]);
        else umma_commit_mc(&empty[s], cmask);
      }
      umma_commit(tmem_full);     // accumulator complete
    }
  } else {
    // ---------------------------------------------------------------- epilogue (warps 2..5)
    const int q = warp & 3;       // TMEM lane quadrant this warp may access
    const int row = m0 + 32 * q + lane;
    if (nkb > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * q) << 16);
    if (EPI == TC_GATED) {
#pragma unroll 1
      for (int c0 = 0; c0 < BN / 2; c0 += 32) {
        uint32_t hv[32], gv[32];
        if (nkb > 0) {
          tmem_ld32(lane_addr + c0, hv);
          tmem_ld32(lane_addr + BN / 2 + c0, gv);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) hv[j] = gv[j] = 0u;
        }
        if (row < p.M) {
          const int col0 = n0 + c0;
          const size_t base = (size_t)row * p.ldc + col0;
#pragma unroll
          for (int j4 = 0; j4 < 32; j4 += 4) {
            float o[4], hh[4], ss[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int col = col0 + j4 + e;
              const float bh = (col < p.gated_O && p.bias0) ? p.bias0[col] : 0.f;
              const float bg = (col < p.gated_O && p.bias1) ? p.bias1[col] : 0.f;
              hh[e] = __uint_as_float(hv[j4 + e]) + bh;
              ss[e] = sigmoidf_(__uint_as_float(gv[j4 + e]) + bg);
              o[e] = hh[e] * ss[e];
            }
            if (p.c_vec && col0 + j4 + 3 < p.gated_O) {
              *reinterpret_cast<float4*>(p.out0 + base + j4) = make_float4(o[0], o[1], o[2], o[3]);
              if (p.out1) *reinterpret_cast<float4*>(p.out1 + base + j4) = make_float4(hh[0], hh[1], hh[2], hh[3]);
              if (p.out2) *reinterpret_cast<float4*>(p.out2 + base + j4) = make_float4(ss[0], ss[1], ss[2], ss[3]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (col0 + j4 + e < p.gated_O) {
                  p.out0[base + j4 + e] = o[e];
                  if (p.out1) p.out1[base + j4 + e] = hh[e];
                  if (p.out2) p.out2[base + j4 + e] = ss[e];
                }
            }
          }
        }
      }
    } else {
      float* dst_base = (EPI == TC_SPLITK) ? p.out0 + (size_t)blockIdx.z * p.M * p.ldc : p.out0;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= p.N) break;   // warp-uniform
        uint32_t v[32];
        if (nkb > 0) {
          tmem_ld32(lane_addr + c0, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        if (row < p.M) {
          const int col0 = n0 + c0;
          float* dst = dst_base + (size_t)row * p.ldc + col0;
#pragma unroll
          for (int j4 = 0; j4 < 32; j4 += 4) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x = __uint_as_float(v[j4 + e]);
              if (EPI == TC_BIAS_ACT) {
                const int col = col0 + j4 + e;
                if (p.bias0 && col < p.N) x += p.bias0[col];
                if (p.act == EXVAE_ACT_SIGMOID) x = sigmoidf_(x);
                else if (p.act == EXVAE_ACT_HARDTANH) x = fminf(fmaxf(x, p.lo), p.hi);
                else if (p.act == EXVAE_ACT_RELU) x = fmaxf(x, 0.f);
              }
              o[e] = x;
            }
            if (p.c_vec && col0 + j4 + 3 < p.N) {
              *reinterpret_cast<float4*>(dst + j4) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (col0 + j4 + e < p.N) dst[j4 + e] = o[e];
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();    // no CTA exits while a peer may still multicast into it / arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// hi = tf32_rna(x), lo = tf32_rna(x - hi): both have their 13 low mantissa bits clear
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ x, size_t n, float* __restrict__ out,
                                                         size_t plane_stride) {
  const size_t n4 = n >> 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const float in[4] = {v.x, v.y, v.z, v.w};
    float h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      uint32_t hb, lb;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(in[e]));
      h[e] = __uint_as_float(hb);
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(in[e] - h[e]));
      l[e] = __uint_as_float(lb);
    }
    reinterpret_cast<float4*>(out)[i] = make_float4(h[0], h[1], h[2], h[3]);
    reinterpret_cast<float4*>(out + plane_stride)[i] = make_float4(l[0], l[1], l[2], l[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (size_t i = n4 << 2; i < n; ++i) {
      uint32_t hb, lb;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x[i]));
      const float hf = __uint_as_float(hb);
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(x[i] - hf));
      out[i] = hf;
      out[plane_stride + i] = __uint_as_float(lb);
    }
}

template <int BN, bool A_MN, bool B_MN, int EPI, int CL>
int launch_cl(const TcGemm& g, const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, cudaStream_t st) {
  constexpr int STAGE = 2 * TBM * TBK * 4 + 2 * BN * TBK * 4;
  constexpr int SMEM = TSTAGES * STAGE + 1024 + 256;
  auto kern = gemm_tf32x3_kernel<BN, A_MN, B_MN, EPI, CL>;
  static bool configured = false;
  if (!configured) {
    EXVAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  const int mtiles = ceil_div(g.M, TBM);
  dim3 grid(EPI == TC_GATED ? ceil_div(g.gated_O, BN / 2) : ceil_div(g.N, BN), ceil_div(mtiles, CL) * CL,
            EPI == TC_SPLITK ? g.splits : 1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(TTHREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = CL;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ma, mb, p);
  return e == cudaSuccess ? EXVAE_OK : (int)e;
}

// cluster size along M: 4 when the row blocks allow it (B-tile traffic / 4), never for split-K (odd tile counts)
template <int BN, bool A_MN, bool B_MN, int EPI>
int launch(const TcGemm& g, const CUtensorMap (&mb)[3], const CUtensorMap& ma, const TcParams& p, int cl,
           cudaStream_t st) {
  if (cl == 4) return launch_cl<BN, A_MN, B_MN, EPI, 4>(g, ma, mb[2], p, st);
  if (cl == 2) return launch_cl<BN, A_MN, B_MN, EPI, 2>(g, ma, mb[1], p, st);
  return launch_cl<BN, A_MN, B_MN, EPI, 1>(g, ma, mb[0], p, st);
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

int tc_split(const float* x, size_t n, float* out, size_t plane_stride, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || (plane_stride & 3))
    return EXVAE_ERR_INVALID_ARG;
  const size_t n4 = (n + 3) / 4;
  const int blocks = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
  split_tf32_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, st>>>(x, n, out, plane_stride);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? EXVAE_OK : (int)e;
}

int tc_cluster_size(const TcGemm& g) {
  static int forced = -1;
  if (forced < 0) {
    const char* env = getenv("EXVAE_GEMM_CLUSTER");
    forced = env ? atoi(env) : 0;
  }
  // Measured on B200 (profiles/r1_gemm_multicast.md): multicasting the B tile over 2 or 4 CTAs does NOT speed
  // the GEMMs up (1.72 / 1.72 / 1.78 ms per step for cluster 1 / 2 / 4): the limiter is the ~40 B/clk each SM can
  // ingest, not the L2 read traffic, and every CTA still receives the full tile.  Default: no cluster.
  if (g.epi == TC_SPLITK || forced <= 1) return 1;
  const int mtiles = ceil_div(g.M, TBM);
  int cl = (mtiles % 4 == 0) ? 4 : (mtiles % 2 == 0) ? 2 : 1;
  return std::min(cl, forced);
}

int tc_gemm_launch(const TcGemm& g, cudaStream_t st) {
  constexpr int BN = 128;
  CUtensorMap ma, mb[3];
  int rc = make_map(&ma, g.a_split, g.a_rows, g.a_cols, g.a_mn ? 32 : TBM, g.a_mn);
  if (rc) return rc;
  const int cl = tc_cluster_size(g);
  // B box rows: whole tile (gated: one half) without clusters, the CTA's 1/cl share with multicast
  const int b_box1 = g.b_mn ? 32 : (g.epi == TC_GATED ? BN / 2 : BN);
  const int idx = cl == 4 ? 2 : cl == 2 ? 1 : 0;
  const int b_box = cl == 1 ? b_box1 : (g.b_mn ? 32 : BN / cl);
  rc = make_map(&mb[idx], g.b_split, g.b_rows, g.b_cols, b_box, g.b_mn);
  if (rc) return rc;
  TcParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.kchunk = g.epi == TC_SPLITK ? g.kchunk : g.K;
  p.gated_O = g.gated_O;
  p.bias0 = g.bias0; p.bias1 = g.bias1; p.out0 = g.out0; p.out1 = g.out1; p.out2 = g.out2; p.ldc = g.ldc;
  p.act = g.act; p.lo = g.lo; p.hi = g.hi;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  p.c_vec = (g.ldc % 4 == 0) && al16(g.out0) && (!g.out1 || al16(g.out1)) && (!g.out2 || al16(g.out2)) &&
            (g.epi != TC_SPLITK || ((size_t)g.M * g.ldc) % 4 == 0);
  if (g.epi == TC_GATED && !g.a_mn && !g.b_mn) return launch<BN, false, false, TC_GATED>(g, mb, ma, p, cl, st);
  if (g.epi == TC_BIAS_ACT && !g.a_mn && !g.b_mn) return launch<BN, false, false, TC_BIAS_ACT>(g, mb, ma, p, cl, st);
  if (g.epi == TC_PLAIN && !g.a_mn && g.b_mn) return launch<BN, false, true, TC_PLAIN>(g, mb, ma, p, cl, st);
  if (g.epi == TC_SPLITK && g.a_mn && g.b_mn) return launch<BN, true, true, TC_SPLITK>(g, mb, ma, p, cl, st);
  return EXVAE_ERR_UNSUPPORTED;
}

}  // namespace exvae
