/* exvae_b200 — C ABI of the B200-native (sm_100a) Exemplar-VAE hot path.
 *
 * The reference (sajadn/Exemplar-VAE) is pure Python/PyTorch and prescribes no FFI; its plug
 * points are Python call sites.  Each entry point below replaces the arithmetic behind one
 * of those call sites (cited as reference `file:line`); the Python package
 * `exemplar_vae_b200` binds them with ctypes and re-exposes the reference's own function
 * and class names.
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless a parameter says "host".  The caller owns every
 *    buffer; the library allocates nothing.  Scratch memory is passed in as `ws`/`ws_bytes`
 *    (query the matching *_workspace_bytes function; 256-byte aligned).
 *  - Tensors are contiguous row-major fp32; indices are int64 (torch.long).
 *  - All work is enqueued on `stream` (a cudaStream_t); no call synchronises the host, so
 *    every call is legal inside CUDA-graph capture.  The library is stateless/re-entrant.
 *  - Return value: 0 = ok; <0 = argument/capability error (EXVAE_ERR_*); >0 = cudaError_t
 *    of the failed launch.  No C++ exception crosses this boundary.
 */
#ifndef EXVAE_B200_H_
#define EXVAE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* exvae_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define EXVAE_API __attribute__((visibility("default")))
#else
#define EXVAE_API
#endif

#define EXVAE_OK 0
#define EXVAE_ERR_INVALID_ARG (-1)
#define EXVAE_ERR_UNSUPPORTED (-2)
#define EXVAE_ERR_WORKSPACE (-3)

#define EXVAE_ABI_VERSION 1

/* activation codes for exvae_linear_* (utils/nn.py:29-41 NonLinear) */
#define EXVAE_ACT_NONE 0
#define EXVAE_ACT_SIGMOID 1
#define EXVAE_ACT_HARDTANH 2
#define EXVAE_ACT_RELU 3

EXVAE_API int exvae_abi_version(void);
EXVAE_API const char* exvae_error_string(int code);
/* host out-params; fails with a cudaError_t if no device is usable */
EXVAE_API int exvae_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------- exemplar prior (K1)
 * Fused replacement of  log_normal_diag_vectorized -> pairwise_distance
 * (utils/distributions.py:12-25), the leave-one-out mask and normaliser
 * (models/BaseModel.py:98-109) and the max-shifted log-sum-exp (models/BaseModel.py:123-125).
 *
 *   z [B,D], mu [C,D], logvar [D] (row 0 of the reference's [C,D] log-variance bank,
 *   BaseModel.py:101), z_idx [B] / mu_idx [C] dataset indices or NULL (either NULL => no mask,
 *   i.e. test mode / no_mask, BaseModel.py:103).
 *
 * fwd writes per-row partial statistics of THIS bank shard: stats[b] = (m2, s2, n_masked, 0)
 * with  sum_n exp(logit[b,n]) = 2^m2 * s2 * exp(c_b)  (c_b is a row constant applied in
 * finalize).  finalize merges the stats of G shards ([G,B,4], G=1 on one GPU) into
 *   log_p[b] = LSE_n logit[b,n] - log(C_total - n_masked_b)          (BaseModel.py:107-108,125)
 * and lse2[b] (base-2 row log-sum, saved for the backward).
 * bwd: given dL/dlog_p returns dz [B,D] (partial over this shard), dmu [C,D] (complete for
 * this shard) and dlogvar [D] (partial over this shard).
 * c_valid (nullable device int32): only the first *c_valid rows of the bank count (the kNN mode selects a
 * data-dependent number of exemplars, models/BaseModel.py:265-266; the bank is then a fixed-capacity [C,D] buffer and
 * the count stays on the device, so the step is graph-capturable).  Rows beyond it contribute nothing (their dmu is
 * 0) and finalize uses *c_valid as the normaliser's exemplar count.
 * D <= 63: dedicated tcgen05 kernels (prior_lse_tc.cu / prior_bwd_tc.cu); D >= 64: the logit tile is a dense
 * z.mu^T contraction and runs through the persistent 3xTF32 tcgen05 GEMM with LSE / weight epilogues (gemm_tc.cu).
 */
EXVAE_API size_t exvae_prior_lse_workspace_bytes(int B, int C, int D);     /* forward + backward */
EXVAE_API size_t exvae_prior_lse_fwd_workspace_bytes(int B, int C, int D); /* forward only (evaluation) */
/* One-GPU use: pass log_p / lse2 ([B] each) and C_total (0 = C): fwd then returns log p(z) itself and no finalize call
 * is needed; for D <= 63 that is ONE kernel from the raw inputs (prior_fused.cu: scaling, tf32 split, mask, log-sum-exp
 * and normaliser on chip).  stats [B,4] is always written (sharded use: exchange it, then call finalize).
 * exvae_prior_lse_fwd_prepares_ws: 1 if fwd leaves `ws` staged for bwd (ws_prepared = 1), 0 if bwd must stage itself. */
EXVAE_API int exvae_prior_lse_fwd_prepares_ws(int B, int C, int D);
EXVAE_API int exvae_prior_lse_fwd(const float* z, const float* mu, const float* logvar, const int64_t* z_idx,
                        const int64_t* mu_idx, int B, int C, int D, const int* c_valid, float* stats /*[B,4]*/,
                        int64_t C_total, float* log_p, float* lse2, void* ws, size_t ws_bytes, exvae_stream_t stream);
EXVAE_API int exvae_prior_lse_finalize(const float* stats /*[G,B,4]*/, int G, const float* z, const float* logvar, int B, int D,
                             int64_t C_total, const int* c_valid, float* log_p /*[B]*/, float* lse2 /*[B]*/,
                             exvae_stream_t stream);
/* ws must be the workspace a fwd call with the same arguments filled (ws_prepared=1) or any
 * workspace of the right size (ws_prepared=0: the bank is re-staged). */
EXVAE_API int exvae_prior_lse_bwd(const float* z, const float* mu, const float* logvar, const int64_t* z_idx,
                        const int64_t* mu_idx, int B, int C, int D, const float* lse2, const float* grad_log_p,
                        float* dz, float* dmu, float* dlogvar, void* ws, size_t ws_bytes, int ws_prepared,
                        const int* c_valid, exvae_stream_t stream);

/* ---------------------------------------------------------------- materialising primitives
 * pairwise_distance (utils/distributions.py:12-18): fp64 inside, fp32 [B,C] out.            */
EXVAE_API int exvae_pairwise_distance(const float* z, const float* means, int B, int C, int D, float* out,
                            exvae_stream_t stream);
/* log_normal_diag_vectorized (utils/distributions.py:21-25); log_var [D]; pair_dist may be NULL */
EXVAE_API int exvae_log_normal_diag_vectorized(const float* x, const float* mean, const float* log_var, int B, int C, int D,
                                     float* log_normal, float* pair_dist, exvae_stream_t stream);
/* log_p_z_exemplar with sum=False (models/BaseModel.py:98-109,126-127): [B,C] matrix with -inf at
 * masked pairs and -log(C - n_masked_b) applied.  row_counts: scratch [B] int32. */
EXVAE_API int exvae_prior_logprob_matrix(const float* z, const float* mu, const float* logvar, const int64_t* z_idx,
                               const int64_t* mu_idx, int B, int C, int D, float* out, int* row_counts,
                               exvae_stream_t stream);

/* ---------------------------------------------------------------- VampPrior (prior == 'vampprior')
 * models/BaseModel.py:84-96,111-128: mixture of C Gaussians with per-component mean AND log-variance
 * (both [C,D], the encodings of the learned pseudo-inputs).  logprob_matrix writes the reference's [B,C]
 * matrix  sum_d log N(z_b | mean_c, exp(logvar_c)) - log C  (log_p_z(sum=False)); lse_fwd also reduces it with the
 * max-shifted log-sum-exp and keeps the matrix (`mat`, caller-owned [B,C]) for the backward, which recomputes
 * the mixture weights from it.                                                                          */
EXVAE_API int exvae_vamp_logprob_matrix(const float* z, const float* mean, const float* logvar, int B, int C, int D,
                                        float* out, exvae_stream_t stream);
EXVAE_API int exvae_vamp_lse_fwd(const float* z, const float* mean, const float* logvar, int B, int C, int D, float* mat,
                                 float* log_p, exvae_stream_t stream);
EXVAE_API int exvae_vamp_lse_bwd(const float* z, const float* mean, const float* logvar, const float* mat,
                                 const float* log_p, const float* grad_log_p, int B, int C, int D, float* dz,
                                 float* dmean, float* dlogvar, exvae_stream_t stream);

/* ---------------------------------------------------------------- kNN exemplar selection (K2)
 * pairwise_distance(z, bank).topk(k, largest=False) (models/BaseModel.py:263-264): fp64
 * expansion distance rounded to fp32, k smallest per row sorted ascending, ties -> lowest
 * position.  metric 1 = direct-difference fp32 Euclidean with sqrt (utils/knn_on_latent.py:4-9).
 * pos_offset is added to every returned position (bank shards).  ONE kernel (knn_fused.cu): distance tiles and a
 * per-row running top-k (k <= 32) stay on chip, the [B,C] matrix is never written; >= 148 CTAs by splitting the
 * bank columns, merged by the last CTA of each row block.                                          */
EXVAE_API size_t exvae_knn_workspace_bytes(int B, int C, int D, int k);
EXVAE_API int exvae_knn_topk(const float* z, const float* bank, int B, int C, int D, int k, int metric, int64_t pos_offset,
                   int64_t* out_idx /*[B,k]*/, float* out_dist /*[B,k]*/, void* ws, size_t ws_bytes,
                   exvae_stream_t stream);
/* merge G per-shard candidate lists [G,B,k] into the global k smallest per row */
EXVAE_API int exvae_knn_merge(const int64_t* idx, const float* dist, int G, int B, int k, int64_t* out_idx, float* out_dist,
                    exvae_stream_t stream);
/* torch.unique(positions) (models/BaseModel.py:265) for positions in [0, range): ascending unique
 * values in out[0..count) (capacity n), count written to *out_count (device int32); out[count..n) repeats
 * out[0] so that the fixed-capacity result only holds valid positions.  flags: scratch [range] int32.   */
EXVAE_API int exvae_unique_positions(const int64_t* pos, int n, int range, int64_t* out, int* out_count, int* flags,
                           exvae_stream_t stream);

/* ---------------------------------------------------------------- row movement
 * dataset.tensors[0][idx] (models/BaseModel.py:247,267), cached_z[idx] (:262) and the cache
 * refresh cached_z[idx] = rows (:261,:269).                                                     */
EXVAE_API int exvae_gather_rows(const float* src, const int64_t* idx, int n_rows, int row_len, float* out,
                      exvae_stream_t stream);
EXVAE_API int exvae_scatter_rows(float* dst, const int64_t* idx, int n_rows, int row_len, const float* src,
                       exvae_stream_t stream);

/* ---------------------------------------------------------------- dense layers (K3)
 * GatedDense (utils/nn.py:44-69): out = (x Wh^T + bh) * sigmoid(x Wg^T + bg).
 * x [R,K], Wh/Wg [O,K], out [R,O].  sig [R,O] = sigmoid(x Wg^T + bg) is saved for the backward (NULL to
 * skip); the backward needs only `out` and `sig`:  dh = dout*sig,  dg = dout*out*(1-sig).
 *
 * Backend: error-compensated 3xTF32 on the tcgen05 tensor cores when the shapes allow TMA (K and
 * the output width multiples of 4, 16-byte aligned pointers) and a forward workspace is given,
 * else the fp32 FMA-pipe GEMM (EXVAE_GEMM=simt forces the latter).  Operands stay plain fp32 in HBM (the
 * kernel splits them into tf32 hi/lo parts in shared memory); the forward workspace of a gated layer holds
 * [Wh ; Wg] as one operand: pass it to the backward (fwd_ws) to reuse it, or NULL.                  */
EXVAE_API size_t exvae_dense_fwd_workspace_bytes(int R, int K, int O, int gated);
EXVAE_API int exvae_gated_dense_fwd(const float* x, const float* Wh, const float* bh, const float* Wg, const float* bg, int R,
                          int K, int O, float* out, float* sig, void* ws, size_t ws_bytes,
                          exvae_stream_t stream);
EXVAE_API size_t exvae_gated_dense_bwd_workspace_bytes(int R, int K, int O);
/* dx may be NULL (first layer: the input is data).  accumulate=1: dW/db are ADDED into the given
 * buffers (fused gradient accumulation straight into the .grad storage) instead of overwritten. */
EXVAE_API int exvae_gated_dense_bwd(const float* x, const float* Wh, const float* Wg, const float* out, const float* sig,
                          const float* dout, int R, int K, int O, float* dx, float* dWh, float* dbh, float* dWg,
                          float* dbg, const void* fwd_ws, size_t fwd_ws_bytes, void* ws, size_t ws_bytes,
                          int accumulate, exvae_stream_t stream);
/* nn.Linear / NonLinear (utils/nn.py:29-41): out = act(x W^T + b); b may be NULL. */
EXVAE_API int exvae_linear_fwd(const float* x, const float* W, const float* b, int R, int K, int O, int act, float lo, float hi,
                     float* out, void* ws, size_t ws_bytes, exvae_stream_t stream);
EXVAE_API size_t exvae_linear_bwd_workspace_bytes(int R, int K, int O);
/* `out` is the forward OUTPUT (post activation); dx / db may be NULL. */
EXVAE_API int exvae_linear_bwd(const float* x, const float* W, const float* out, const float* dout, int R, int K, int O, int act,
                     float lo, float hi, float* dx, float* dW, float* db, const void* fwd_ws, size_t fwd_ws_bytes,
                     void* ws, size_t ws_bytes, int accumulate, exvae_stream_t stream);
/* 1 = tcgen05 3xTF32 backend active on the current device, 0 = fp32 FMA-pipe backend */
EXVAE_API int exvae_gemm_backend(void);
/* Debug / profiling aid (tools/gemm_trace.py, tools/prior_trace.py): when buf != NULL every CTA of the following
 * tensor-core launches writes 8 uint64 of %globaltimer stamps into the launch's segment of buf (GEMM: 160 CTAs x 8
 * words per launch: {start, -, tiles done, -, -, -, end, SM id}; prior backward: 400 x 8 words per pass, see
 * prior_bwd_tc.cu); NULL switches tracing off.  Keep the stamps out of hot loops: a %globaltimer read is slow. */
EXVAE_API int exvae_gemm_set_trace(uint64_t* buf);

/* Deferred finish of the weight gradients (utils/training.py:38-39: `loss.backward(); optimizer.step()` -- nothing reads
 * a parameter gradient between the two).  With on != 0 the *_dense_bwd / linear_bwd calls of LARGE layers (R > 4096,
 * accumulate != 0) queue their last pass (split-K reduction of dW + bias column sums, added into dW / db) on the
 * library's side stream and return WITHOUT joining it; exvae_dense_bwd_flush(stream) makes `stream` wait for every
 * queued finish.  The caller keeps each call's workspace alive until the flush and flushes before anything reads the
 * gradients.  Returns the previous setting (defer_finish) / 0 (flush).  Graph-capturable (event fork / join). */
EXVAE_API int exvae_dense_bwd_defer_finish(int on);
EXVAE_API int exvae_dense_bwd_flush(exvae_stream_t stream);

/* ---------------------------------------------------------------- convolution support (K4)
 * GatedConv2d / Conv2d (utils/nn.py:72-114) and the weight-normed conv / ELU / Upsample blocks of
 * models/fully_conv.py:12-81 run as  im2col -> dense-layer GEMM (K3, fused gate/bias/activation)
 * -> col2im.  Activations are NHWC; col is [N*OH*OW, kh*kw*C] with the channel fastest, where
 * OH = (H + 2*pad - kh)/stride + 1 (same for OW).  col2im is the exact adjoint of im2col.           */
EXVAE_API int exvae_im2col_nhwc(const float* x, int N, int H, int W, int C, int kh, int kw, int stride, int pad,
                                float* col, exvae_stream_t stream);
EXVAE_API int exvae_col2im_nhwc(const float* dcol, int N, int H, int W, int C, int kh, int kw, int stride, int pad,
                                float* dx, exvae_stream_t stream);
/* torch.nn.ELU (alpha = 1); the backward takes the forward OUTPUT y */
EXVAE_API int exvae_elu_fwd(const float* x, int64_t n, float* y, exvae_stream_t stream);
EXVAE_API int exvae_elu_bwd(const float* y, const float* dy, int64_t n, float* dx, exvae_stream_t stream);
/* ---------------------------------------------------------------- convolution (K4)
 * GatedConv2d / Conv2d (utils/nn.py:72-114) and the weight-normed convolutions of models/fully_conv.py, NHWC.
 * Layers with >= 16 input channels run as an IMPLICIT GEMM on tcgen05 (no patch matrix in HBM: one 4-D TMA box of the
 * activation tensor per filter tap and 32-channel chunk, zero padding by TMA's out-of-bounds fill); so does the input
 * gradient of stride-1 layers (convolution of the pre-activation gradient with the flipped filters).  Layers with
 * fewer input channels, the weight gradient and the input gradient of stride-2 layers use a patch matrix whose rows
 * are padded to 4 floats (every GEMM on tcgen05).
 * exvae_conv_plan: plan[0..8] = {OH, OW, implicit, cpad, Kp, Kpc, dx_implicit, cpad_dx, Kp_dx}; ncat = 2*O (gated) or O.
 *   forward operand  wpk [ncat][Kp]   = exvae_conv_pack_weight(mode 0, cpad, Kp)  (gated: h rows then g rows)
 *   backward operand wbw              = dx_implicit ? mode 1, cpad_dx, Kp_dx, rows = Cin : mode 0, cpad = Cin, Kpc
 * pack_weight: w [Cout][Cin][KH][KW] (the reference's nn.Conv2d layout) -> rows [row_off, row_off+Cout) (mode 0) or
 * columns row_off.. of every tap (mode 1) of `out` [rows_total][Kp]; zero_first clears `out` (padding must be 0). */
EXVAE_API int exvae_conv_plan(int N, int H, int W, int Cin, int KH, int KW, int stride, int pad, int ncat, int* plan);
EXVAE_API int exvae_conv_pack_weight(const float* w, int Cout, int Cin, int KH, int KW, int mode, int cpad, int Kp,
                                     int row_off, int zero_first, int rows_total, float* out, exvae_stream_t stream);
EXVAE_API size_t exvae_conv2d_fwd_workspace_bytes(int N, int H, int W, int Cin, int KH, int KW, int stride, int pad,
                                                  int ncat);
/* out [N,OH,OW,O] = act(conv(x, W0) + b0)                      (gated = 0; act as in exvae_linear_fwd)
 *                 = (conv(x,W0)+b0) * sigmoid(conv(x,W1)+b1)   (gated = 1; sig [N,OH,OW,O] saved for the backward) */
EXVAE_API int exvae_conv2d_fwd(const float* x, const float* wpk, const float* b0, const float* b1, int N, int H, int W,
                               int Cin, int KH, int KW, int stride, int pad, int O, int gated, int act, float lo,
                               float hi, float* out, float* sig, void* ws, size_t ws_bytes, exvae_stream_t stream);
EXVAE_API size_t exvae_conv2d_bwd_workspace_bytes(int N, int H, int W, int Cin, int KH, int KW, int stride, int pad, int O,
                                                  int gated);
/* dx [N,H,W,Cin] (nullable), dW0/dW1 [O][Cin][KH][KW], db0/db1 [O] (nullable); accumulate != 0 adds into them. */
EXVAE_API int exvae_conv2d_bwd(const float* x, const float* wbw, const float* out, const float* sig, const float* dout,
                               int N, int H, int W, int Cin, int KH, int KW, int stride, int pad, int O, int gated,
                               int act, float lo, float hi, float* dx, float* dW0, float* db0, float* dW1, float* db1,
                               void* ws, size_t ws_bytes, int accumulate, exvae_stream_t stream);
/* torch.nn.utils.weight_norm (models/fully_conv.py:17): w[r,:] = g[r] * v[r,:] / ||v[r,:]||_2 for R rows of length K;
 * backward: dv, dg from dw (accumulate != 0 adds into them). */
EXVAE_API int exvae_weight_norm_fwd(const float* v, const float* g, int R, int K, float* w, exvae_stream_t stream);
EXVAE_API int exvae_weight_norm_bwd(const float* v, const float* g, const float* dw, int R, int K, float* dv, float* dg,
                                    int accumulate, exvae_stream_t stream);
/* nn.Upsample(scale_factor=2) (nearest), NHWC: y [N,2H,2W,C]; backward sums each 2x2 block */
EXVAE_API int exvae_upsample2x_nhwc_fwd(const float* x, int N, int H, int W, int C, float* y, exvae_stream_t stream);
EXVAE_API int exvae_upsample2x_nhwc_bwd(const float* dy, int N, int H, int W, int C, float* dx, exvae_stream_t stream);

/* ---------------------------------------------------------------- element-wise pieces
 * reparameterize (models/BaseModel.py:79-82): z = mu + exp(0.5 logvar) * eps (eps injected). */
EXVAE_API int exvae_reparameterize_fwd(const float* mu, const float* logvar, const float* eps, int64_t n, float* z,
                             exvae_stream_t stream);
EXVAE_API int exvae_reparameterize_bwd(const float* logvar, const float* eps, const float* dz, int64_t n, float* dmu,
                             float* dlogvar, exvae_stream_t stream);
/* reparameterize + log q(z|x) in one pass (models/BaseModel.py:79-82 followed by utils/distributions.py:28-33 as
 * models/AbsModel.py:18 / AbsHModel.py:21,27 call them): z [B,D] and logq [B] = log_normal_diag(z, mu, logvar, dim=1),
 * bit-identical to the two separate calls.  Backward: dz = gradient reaching z from its other consumers (nullable),
 * dlogq [B] (nullable) -> dmu, dlogvar (total derivatives through z). */
EXVAE_API int exvae_reparam_logq_fwd(const float* mu, const float* logvar, const float* eps, int B, int D, float* z,
                                     float* logq, exvae_stream_t stream);
EXVAE_API int exvae_reparam_logq_bwd(const float* mu, const float* logvar, const float* eps, const float* z,
                                     const float* dz, const float* dlogq, int B, int D, float* dmu, float* dlogvar,
                                     exvae_stream_t stream);
/* prior_log_variance [1] broadcast to a [n] row (models/BaseModel.py:212-214) and the sum of its gradient
 * (accumulate != 0: out[0] += sum, for in-place accumulation into the parameter's .grad). */
EXVAE_API int exvae_bcast_scalar(const float* src, int n, float* out, exvae_stream_t stream);
EXVAE_API int exvae_sum_to_scalar(const float* src, int n, float* out, int accumulate, exvae_stream_t stream);
/* torch.cat((a, b), 1) (models/AbsHModel.py:55,83): out [R, Ka+Kb]; backward splits dout (da / db nullable). */
EXVAE_API int exvae_concat_cols_fwd(const float* a, const float* b, int64_t R, int Ka, int Kb, float* out,
                                    exvae_stream_t stream);
EXVAE_API int exvae_concat_cols_bwd(const float* dout, int64_t R, int Ka, int Kb, float* da, float* db,
                                    exvae_stream_t stream);
/* out[i] = src[idx[i]] for int64 vectors: exemplars_indices[nearest] (models/BaseModel.py:266) */
EXVAE_API int exvae_gather_index(const int64_t* src, const int64_t* idx, int n, int64_t* out, exvae_stream_t stream);
/* zero `bytes` bytes (cudaMemsetAsync): the one gradient-buffer reset of a training step. */
EXVAE_API int exvae_zero(void* ptr, size_t bytes, exvae_stream_t stream);
/* log_normal_diag(x, mean, log_var, dim=1) (utils/distributions.py:28-33): out [B]. */
EXVAE_API int exvae_log_normal_diag_fwd(const float* x, const float* mean, const float* logvar, int B, int D, float* out,
                              exvae_stream_t stream);
EXVAE_API int exvae_log_normal_diag_bwd(const float* x, const float* mean, const float* logvar, const float* dout, int B, int D,
                              float* dx, float* dmean, float* dlogvar, exvae_stream_t stream);
/* log_normal_standard(x, dim=1) (utils/distributions.py:36-41) */
EXVAE_API int exvae_log_normal_standard_fwd(const float* x, int B, int D, float* out, exvae_stream_t stream);
EXVAE_API int exvae_log_normal_standard_bwd(const float* x, const float* dout, int B, int D, float* dx, exvae_stream_t stream);
/* log_bernoulli(x, mean, dim=1) (utils/distributions.py:44-51), probs clamped to [1e-5, 1-1e-5] */
EXVAE_API int exvae_log_bernoulli_fwd(const float* x, const float* mean, int B, int P, float* out, exvae_stream_t stream);
EXVAE_API int exvae_log_bernoulli_bwd(const float* x, const float* mean, const float* dout, int B, int P, float* dmean,
                            exvae_stream_t stream);
/* log_logistic_256(x, mean, logvar, dim=1) (utils/distributions.py:54-66) */
EXVAE_API int exvae_log_logistic256_fwd(const float* x, const float* mean, const float* logvar, int B, int P, float* out,
                              exvae_stream_t stream);
EXVAE_API int exvae_log_logistic256_bwd(const float* x, const float* mean, const float* logvar, const float* dout, int B, int P,
                              float* dmean, float* dlogvar, exvae_stream_t stream);
/* loss = mean(-RE + beta*KL), RE.mean, KL.mean (models/BaseModel.py:71-75); out3 = {loss, RE, KL};
 * average=0 writes per-sample loss into loss_b [B] instead (out3 may be NULL).  beta_dev (nullable): a device
 * scalar that overrides `beta`, so a captured CUDA graph follows the warm-up schedule of utils/training.py:5-12. */
EXVAE_API int exvae_elbo_reduce(const float* RE, const float* KL, int B, float beta, const float* beta_dev, int average,
                                float* out3, float* loss_b, exvae_stream_t stream);

/* out[i] = sum_j c[j] * x_j[i] over up to 4 vectors (NULL x_j are skipped): the KL assembly
 * -(log_p_z1 + log_p_z2 - log_q_z1 - log_q_z2) of models/AbsModel.py:19, models/AbsHModel.py:29. */
EXVAE_API int exvae_lincomb4(const float* x0, const float* x1, const float* x2, const float* x3, float c0, float c1,
                             float c2, float c3, int64_t n, float* out, exvae_stream_t stream);
/* backward of exvae_elbo_reduce: g3 = upstream grads of {loss, RE, KL} (average=1) or g_loss_b [B]
 * (average=0, g3 ignored); writes dRE [B], dKL [B]. */
EXVAE_API int exvae_elbo_reduce_bwd(const float* g3, const float* g_loss_b, int B, float beta, const float* beta_dev,
                                    int average, float* dRE, float* dKL, exvae_stream_t stream);

/* ---------------------------------------------------------------- NVLink / NVSwitch exchanges (multi-GPU)
 * The reference is single-device.  The exchanges of the range-sharded step run over SYMMETRIC memory (one buffer per
 * rank with the same layout, mapped into every peer and into one NVSwitch multicast address; set up by the host with
 * torch.distributed._symmetric_memory): mc_ptr arguments are MULTICAST addresses (multimem.st = one store lands in all
 * GPUs, multimem.ld_reduce = the sum over all GPUs computed in the switch), signal_pads_dev is the device array of
 * the ranks' signal pads ([channel][world] uint32 flags) used by the in-kernel barrier.  Each launch uses channels
 * [channel0, channel0 + blocks), blocks <= max_blocks; launches that may run concurrently need disjoint channels.
 *   allreduce      in place over n floats of the buffer (n % (4*world) == 0): buffer = scale * sum over ranks
 *   allgather      src (local, n floats) -> slot `rank` of the [world][n] region mc_dst, on every rank
 *   reduce_scatter out (local, n floats) = sum over ranks of slice `rank` of the [world][n] region mc_src            */
EXVAE_API int exvae_mc_allreduce(float* mc_ptr, void* signal_pads_dev, int64_t n, int rank, int world, int channel0,
                                 int max_blocks, float scale, exvae_stream_t stream);
EXVAE_API int exvae_mc_allgather(const float* src, float* mc_dst, void* signal_pads_dev, int64_t n, int rank, int world,
                                 int channel0, int max_blocks, exvae_stream_t stream);
EXVAE_API int exvae_mc_reduce_scatter(const float* mc_src, float* out, void* signal_pads_dev, int64_t n, int rank,
                                      int world, int channel0, int max_blocks, exvae_stream_t stream);

/* ---------------------------------------------------------------- counter-based RNG (Philox4x32-10)
 * Device-side replacements for torch.bernoulli (utils/training.py:31), torch.randint
 * (models/BaseModel.py:245,257) and normal_() (models/BaseModel.py:81).  `counter` is a device
 * uint64 that the kernel reads as the stream offset, so graph replays draw fresh numbers once it has been
 * bumped: by exvae_rng_advance, or by the drawing kernel itself when advance != 0 (then `counter` must hold TWO
 * uint64: counter[0] the offset, counter[1] a zero-initialised ticket word the kernel uses to find its last block,
 * which adds 1 to counter[0] after every block has read it).  counter may be NULL (offset 0, no advance).   */
EXVAE_API int exvae_rng_bernoulli(const float* p, int64_t n, uint64_t seed, uint64_t* counter, uint64_t subseq, int advance,
                        float* out, exvae_stream_t stream);
EXVAE_API int exvae_rng_normal(int64_t n, uint64_t seed, uint64_t* counter, uint64_t subseq, int advance, float* out,
                     exvae_stream_t stream);
EXVAE_API int exvae_rng_randint(int64_t low, int64_t high, int64_t n, uint64_t seed, uint64_t* counter, uint64_t subseq,
                      int advance, int64_t* out, exvae_stream_t stream);
EXVAE_API int exvae_rng_advance(uint64_t* counter, uint64_t by, exvae_stream_t stream);

/* ---------------------------------------------------------------- AdamNormGrad (utils/optimizer.py:32-80)
 * One fused multi-tensor step: per-tensor g / (||g||_2 + 1e-7), Adam moments, bias correction
 * from the device step counter (incremented by the call), parameter update.
 * table: device array of n_tensors records {param*, grad*, exp_avg*, exp_avg_sq*, numel} laid out
 * as 5 x int64 per tensor.  norms: scratch [16 * n_tensors + 1] fp32 (16 partial squared
 * norms per tensor + the step size).  step: device int64[1].                                                   */
EXVAE_API int exvae_adam_normgrad_step(const int64_t* table, int n_tensors, int64_t max_numel, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int64_t* step, float* norms,
                             exvae_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EXVAE_B200_H_ */
