// Internal interface of the tensor-core exemplar-prior forward (prior_lse_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace exvae {

struct PriorTcArgs {
  const float* zp;   // [2, Bpad, KP]
  const float* mp;   // [2, Cpad, KP]
  int Bpad, Cpad, KP, B, C;
  const int* mcnt;   // NULL => no mask
  const int* mlist;
  const int64_t* cidx;
  const int64_t* z_idx;
  float* part;       // [Bpad, nsplit, 4]
};

// K1 backward on the tensor cores (prior_bwd_tc.cu).  All pointers are workspace arrays staged by
// prior_stage_kernel / prior_bwd_prep (padded to Bpad / Cpad rows).
struct PriorBwdTcArgs {
  const float* zp;   // [2, Bpad, KP]  hi/lo planes of (zs*log2e | 1 | 0..)
  const float* mp;   // [2, Cpad, KP]  hi/lo planes of (ms | nb2 | 0..)
  const float* zsT;  // [2, NG, Bpad]  hi/lo planes of zs^T (rows d >= D are zero)
  const float* msT;  // [2, NG, Cpad]  hi/lo planes of ms^T
  const float* zs;   // [Bpad, LD]
  const float* ms;   // [Cpad, LD]
  const float* glp;  // [Bpad] upstream gradient (0 in the padding)
  const float* lsp;  // [Bpad] base-2 row log-sum (+inf in the padding)
  const int64_t* zip;   // [Bpad] dataset index of each z row (INT64_MIN in the padding), NULL => no mask
  const int64_t* cidx;  // [Cpad] dataset index of each exemplar (INT64_MIN in the padding)
  const float* isig;    // [LD] 1/sigma
  int Bpad, Cpad, KP, NG, LD, B, C, D;
  float* dzs_part;      // [nsplit, Bpad, LD]   (out) per-split W.ms
  float* rowsum_part;   // [nsplit, Bpad]       (out) per-split row sums of W
  float* dmu;           // [C, D]               (out)
  float* coldot_part;   // [Cpad/128, LD]       (out) per column tile sum_n dms[n,d]*ms[n,d]
  float* gcol_part;     // [rsplit, Cpad, NG]   (scratch) pass-2 partials of W^T.zs when the row blocks are split
  float* tot_part;      // [rsplit, Cpad]       (scratch) pass-2 partial column sums of W
};
// row-block splits of pass 2 for this geometry (1 = one CTA per exemplar tile walks all row blocks)
int prior_bwd_pass2_splits(int Bpad, int Cpad);
// launches both passes; *nsplit_out = number of dzs/rowsum partials per row, *ntile_out = column tiles of coldot_part
int prior_bwd_tc_launch(const PriorBwdTcArgs& a, int* nsplit_out, int* ntile_out, cudaStream_t st);
// transposed hi/lo planes + padded per-row arrays for the backward
int prior_bwd_prep_launch(const float* zs, const float* ms, const float* g, const float* lse2, const int64_t* z_idx,
                          int B, int C, int D, int LD, int Bpad, int Cpad, int NG, float* zsT, float* msT, float* glp,
                          float* lsp, int64_t* zip, cudaStream_t st);

bool prior_tc_enabled();
// K1 forward as ONE kernel from the raw inputs (prior_fused.cu; D <= 63, <= 64 row blocks).  ws: prior_fused_ws_bytes.
bool prior_fused_ok(int D);
size_t prior_fused_ws_bytes(int B, int C);
// staged operands the backward reads (workspace arrays of prior_ws_layout); null pointer to the struct = forward only
struct PriorFusedStage {
  float* zs; float* ms; float* zp; float* mp; int64_t* cidx; float* isig;
  int LD, Bpad, Cpad;
};
int prior_fused_fwd_launch(const float* z, const float* mu, const float* logvar, const int64_t* z_idx, const int64_t* mu_idx,
                           const int* c_valid, int B, int C, int D, float c_total, float* stats, float* log_p, float* lse2,
                           void* ws, const PriorFusedStage* sg, cudaStream_t st);
// launches the kernel; *nsplit_out = number of partials written per row
int prior_fwd_tc_launch(const PriorTcArgs& a, int* nsplit_out, cudaStream_t st);

}  // namespace exvae
