/*
 * bdlru.h — C ABI of libbdlru.so: the B200 (sm_100a) implementation of RecBLR's BD-LRU hot path.
 *
 * Drop-in boundary.  Every entry point takes raw DEVICE pointers, explicit sizes and
 * element strides, a dtype tag and a cudaStream_t (passed as void*); it enqueues work on that stream
 * and returns without synchronising.  The caller owns every buffer, including outputs and workspaces
 * (sizes from the *_workspace_bytes queries).  The library keeps no tensors and never throws across
 * the ABI: functions return 0 on success or a BDLRU_ERR_* code, with a thread-local message available
 * from bdlru_last_error().  There is no CPU fallback anywhere behind this header.
 *
 * Reference interfaces replaced (paths relative to /root/reference):
 *   bdlru_scan_fwd / _bwd            parallel_scan.py:85-95 (Scan.forward + forward_scan 44-60)
 *                                    parallel_scan.py:98-114 (Scan.backward + backward_scan 63-80)
 *   bdlru_gated_scan_fwd / _bwd      RecBLR.py:197-200 (gate math + both transposes + parallel_scan)
 *                                    and the left-pad of RecBLR.py:177-179,203-204 (as h0 / dh0)
 *   bdlru_conv1d_fwd / _bwd          causal_conv1d_fn call site RecBLR.py:188-193 (fallback line 185)
 *   bdlru_embed_ln_fwd / _bwd        RecBLR.py:76-78 (embedding gather -> dropout -> LayerNorm)
 *   bdlru_add_ln_fwd / _bwd          RecBLR.py:142, 221-225 (dropout -> + residual -> LayerNorm)
 *   bdlru_silu_dropout_fwd / _bwd    RecBLR.py:219-221 (FFN activation + dropout)
 *   bdlru_colsum                     bias gradients of the nn.Linear layers with bias (RecBLR.py:165, 213-214)
 *   bdlru_fullsort_topk              RecBLR.py:114-122 + RecBole mask/top-k ([upstream])
 *   bdlru_fullsort_ce_fwd / _bwd     RecBLR.py:99-103 (logits GEMM + nn.CrossEntropyLoss, mean)
 */
#ifndef BDLRU_H_
#define BDLRU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BDLRU_OK 0
#define BDLRU_ERR_INVALID 1     /* bad argument (shape, stride, alignment, null pointer) */
#define BDLRU_ERR_CUDA 2        /* a CUDA runtime / driver call failed */
#define BDLRU_ERR_UNSUPPORTED 3 /* valid request this build does not implement */

#define BDLRU_F32 0
#define BDLRU_BF16 1

/* ABI version (bumped on any signature change) and last error text of the calling thread. */
int bdlru_version(void);
const char* bdlru_last_error(void);
/* Number of kernel launches enqueued by this library since process start (bench.py's gpu_launches). */
uint64_t bdlru_launch_count(void);
/* "src=<sha256 of csrc/ + include/ + flags>[ tuning]": which sources this binary was built from, and whether it is a
 * -DBDLRU_TUNING build (the only kind that reads BDLRU_* environment switches).  bench.py refuses tuning builds. */
const char* bdlru_build_info(void);

/* ---------------------------------------------------------------------------------------------
 * S0 — raw first-order scan on [B, C, T] contiguous fp32, T contiguous.
 * Replaces parallel_scan.py:85-95 / 98-114.  h_t = gates_t * h_{t-1} + tokens_t, h_{-1} = 0.
 * Any T >= 1 (the reference's power-of-two restriction, parallel_scan.py:48, is gone).
 *   bwd: d_tokens_t = g_t + gates_{t+1} * d_tokens_{t+1};  d_gates_t = states_{t-1} * d_tokens_t.
 * ------------------------------------------------------------------------------------------- */
int bdlru_scan_fwd(const float* gates, const float* tokens, float* states,
                   int64_t rows /* B*C */, int64_t T, void* stream);
int bdlru_scan_bwd(const float* gates, const float* states, const float* grad_out,
                   float* d_gates, float* d_tokens, int64_t rows, int64_t T, void* stream);

/* ---------------------------------------------------------------------------------------------
 * S1 — fused gate math + scan on channel-last [B, T, C] views (RecBLR.py:197-200 without the
 * transposes, the pad copy or any [B,T,C] intermediate).
 *   a_t = exp(-softplus(Lambda) * sigmoid(r_t));  b_t = sqrt(1 - a_t^2 + 1e-8) * sigmoid(i_t) * xp_t
 *   h_t = a_t * h_{t-1} + b_t,  h_{-1} = h0 (NULL -> 0).   Optional fused z-gate (RecBLR.py:206):
 *   y_t = silu(z_t) * h_t  when z != NULL.
 * Tensors are described by (pointer, batch stride, row stride) in ELEMENTS; the channel stride is 1.
 * xp/r/i/z/h/y share `dtype` (BDLRU_F32 or BDLRU_BF16); Lambda, h0, dLambda, dh0 are fp32; the scan
 * state is fp32.  h0_bstride is 0 for a batch-independent h0[C] (the left-pad leak) or C for h0[B,C].
 * C % 4 == 0; pointers and strides must keep 4-element vectors aligned.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* ptr;
  int64_t bstride; /* elements between batches */
  int64_t rstride; /* elements between time steps */
} bdlru_view;

int bdlru_gated_scan_fwd(bdlru_view xp, bdlru_view r, bdlru_view i, const float* Lambda,
                         const float* h0, int64_t h0_bstride,
                         bdlru_view z /* ptr NULL: no z-gate */,
                         bdlru_view h /* out */, bdlru_view y /* out, used iff z.ptr */,
                         int B, int T, int C, int dtype, void* stream);

/* Backward.  grad is dL/dh (z.ptr == NULL) or dL/dy (z.ptr != NULL; then dz is written too and h must be
 * the forward's h).  Outputs dxp, dr, di (dz) in `dtype`; dLambda[C], dh0 ([C] if h0_bstride == 0 else
 * [B,C]; may be NULL) in fp32, OVERWRITTEN (not accumulated), deterministic (two-pass reduction).
 * workspace: bdlru_gated_scan_bwd_workspace_bytes(B, T, C) bytes of device memory. */
size_t bdlru_gated_scan_bwd_workspace_bytes(int B, int T, int C);
int bdlru_gated_scan_bwd(bdlru_view xp, bdlru_view r, bdlru_view i, const float* Lambda,
                         const float* h0, int64_t h0_bstride, bdlru_view z,
                         bdlru_view h, bdlru_view grad,
                         bdlru_view dxp, bdlru_view dr, bdlru_view di, bdlru_view dz,
                         float* dLambda, float* dh0, void* workspace, size_t workspace_bytes,
                         int B, int T, int C, int dtype, void* stream);

/* The left-pad quirk as an initial state (RecBLR.py:177-199): with s = silu(conv_bias), g = gates_w s + gates_b
 * ([2C]: recurrence | input), a = exp(-softplus(Lambda) * sigmoid(g_rec)), b' = sqrt(1 - a^2 + 1e-8) * sigmoid(g_in) * s:
 *   h0 = b' * sum_{k < pad_len} a^k      (fp32 [C], batch independent; pad_len = 2^ceil(log2 T) - T >= 1)
 * `saved` is a [5*C] fp32 scratch written by fwd and read by bwd.  bwd OVERWRITES dconv_bias [C], dgates_w [2C, C],
 * dgates_b [2C], dLambda [C] with the gradient contributions of dh0. */
int bdlru_phantom_h0_fwd(const float* conv_bias, const float* gates_w, const float* gates_b, const float* Lambda, int C,
                         int pad_len, float* h0, float* saved, void* stream);
int bdlru_phantom_h0_bwd(const float* conv_bias, const float* gates_w, const float* Lambda, const float* saved,
                         const float* dh0, int C, int pad_len, float* dconv_bias, float* dgates_w, float* dgates_b,
                         float* dLambda, void* stream);

/* Channel-last raw scan (no gate math): h_t = a_t * h_{t-1} + b_t on [B, T, C] views, fp32 or bf16 I/O.
 * Same tiling as S1; used when gates are produced elsewhere.  bwd writes da, db (and dh0 like S1). */
int bdlru_scan_cl_fwd(bdlru_view a, bdlru_view b, const float* h0, int64_t h0_bstride, bdlru_view h,
                      int B, int T, int C, int dtype, void* stream);
int bdlru_scan_cl_bwd(bdlru_view a, const float* h0, int64_t h0_bstride, bdlru_view h, bdlru_view grad,
                      bdlru_view da, bdlru_view db, float* dh0, void* workspace, size_t workspace_bytes,
                      int B, int T, int C, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Causal depthwise conv1d (+bias, +SiLU) on channel-last [B, T, C] views.
 * Replaces causal_conv1d_fn(x=[B,C,T] channel-last strides, weight=[C,W], bias=[C], activation)
 * (RecBLR.py:188-193; semantics pinned by the fallback RecBLR.py:185).
 *   y_t = act(bias + sum_{j<W} weight[c][j] * x_{t-(W-1)+j}),  x_{<0} = 0,  W in [1, 4].
 * weight/bias/dweight/dbias are fp32; bias may be NULL; silu != 0 applies SiLU.
 * ------------------------------------------------------------------------------------------- */
int bdlru_conv1d_fwd(bdlru_view x, const float* weight, const float* bias, bdlru_view y,
                     int B, int T, int C, int W, int silu, int dtype, void* stream);
size_t bdlru_conv1d_bwd_workspace_bytes(int B, int T, int C, int W);
int bdlru_conv1d_bwd(bdlru_view x, const float* weight, const float* bias, bdlru_view grad_y,
                     bdlru_view dx, float* dweight /* [C,W] */, float* dbias /* [C] or NULL */,
                     void* workspace, size_t workspace_bytes,
                     int B, int T, int C, int W, int silu, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Front end: out[n,:] = LayerNorm(dropout(table[ids[n],:])) * gamma + beta   (RecBLR.py:76-78).
 * ids int64 [n_tokens]; table fp32 or bf16 [n_items, D] (dtype); out / grad_out in `out_dtype` (= dtype, or bf16 for an
 * fp32 table: mixed-precision training keeps fp32 master rows and bf16 activations); gamma/beta fp32.
 * Dropout (p in [0,1)) uses a counter-based generator keyed by (seed, token, channel); the mask is
 * recomputed in the backward from the same seed.  seed_device (may be NULL) points to a DEVICE uint64 added to
 * `seed` inside the kernel, so a step captured in a CUDA graph draws a fresh mask on every replay.  rstd/mean (fp32 [n_tokens]) are saved for backward.
 * bwd scatter-adds into dtable (fp32 [n_items, D], NOT zeroed here) skipping ids == padding_idx
 * (pass -1 for none), and writes dgamma/dbeta.
 * ------------------------------------------------------------------------------------------- */
int bdlru_embed_ln_fwd(const int64_t* ids, const void* table, const float* gamma, const float* beta,
                       void* out, float* mean, float* rstd, int64_t n_tokens, int64_t n_items, int D,
                       float eps, float dropout_p, uint64_t seed, const uint64_t* seed_device, int dtype, int out_dtype,
                       void* stream);
size_t bdlru_embed_ln_bwd_workspace_bytes(int64_t n_tokens, int D);
int bdlru_embed_ln_bwd(const int64_t* ids, const void* table, const float* gamma, const void* grad_out,
                       const float* mean, const float* rstd, float* dtable, float* dgamma, float* dbeta,
                       void* workspace, size_t workspace_bytes, int64_t n_tokens, int64_t n_items, int D,
                       float dropout_p, uint64_t seed, const uint64_t* seed_device, int64_t padding_idx, int dtype,
                       int out_dtype, void* stream);
/* Row-sharded item table (SURVEY §8e "input gather with a sharded tied table"): the same backward, but instead of
 * scattering it writes the per-token row gradients drows [n_tokens, D] in `out_dtype` (zeros at ids == padding_idx) —
 * the payload of the exchange to the row owners — and bdlru_scatter_add_rows is the owner-side half:
 * dst[(id - row_lo) * D + :] += rows[n, :] for every token with row_lo <= id < row_hi and id != padding_idx
 * (dst = this rank's fp32 gradient shard, NOT zeroed here; vector red.global.add). */
int bdlru_embed_ln_bwd_rows(const int64_t* ids, const void* table, const float* gamma, const void* grad_out,
                            const float* mean, const float* rstd, void* drows, float* dgamma, float* dbeta,
                            void* workspace, size_t workspace_bytes, int64_t n_tokens, int64_t n_items, int D,
                            float dropout_p, uint64_t seed, const uint64_t* seed_device, int64_t padding_idx, int dtype,
                            int out_dtype, void* stream);
int bdlru_scatter_add_rows(const int64_t* ids, const void* rows, int64_t n_tokens, int D, int rows_dtype,
                           int64_t row_lo, int64_t row_hi, int64_t padding_idx, float* dst, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Residual epilogue: out[n,:] = LayerNorm(dropout(x[n,:]) + residual[n,:]) * gamma + beta
 * (RecurrentLayer.forward RecBLR.py:142 and FeedForward.forward RecBLR.py:221-225).  x, residual, out, grad_out, dx,
 * dresidual are contiguous [n_rows, D] in `dtype`; gamma/beta/dgamma/dbeta/mean/rstd fp32.  Same dropout generator and
 * seed_device convention as bdlru_embed_ln_*.  bwd: dresidual = d(sum), dx = dresidual * mask; dx may alias
 * dresidual when dropout_p == 0.  Alignment: fp32 rows 16 bytes; bf16 rows 8 bytes — with D % 8 == 0 and every row
 * pointer 16-byte aligned the kernels move 16 bytes per lane, otherwise 8; the dropout mask is the same either way.
 * bwd workspace: bdlru_add_ln_bwd_workspace_bytes(n_rows, D).
 * ------------------------------------------------------------------------------------------- */
int bdlru_add_ln_fwd(const void* x, const void* residual, const float* gamma, const float* beta, void* out,
                     float* mean, float* rstd, int64_t n_rows, int D, float eps, float dropout_p, uint64_t seed,
                     const uint64_t* seed_device, int dtype, void* stream);
size_t bdlru_add_ln_bwd_workspace_bytes(int64_t n_rows, int D);
int bdlru_add_ln_bwd(const void* x, const void* residual, const float* gamma, const void* grad_out, const float* mean,
                     const float* rstd, void* dx, void* dresidual, float* dgamma, float* dbeta, void* workspace,
                     size_t workspace_bytes, int64_t n_rows, int D, float dropout_p, uint64_t seed,
                     const uint64_t* seed_device, int dtype, void* stream);

/* FeedForward activation (RecBLR.py:219-221): out = dropout(silu(x)) over n contiguous elements (n % 8 == 0, 16-byte aligned), and its
 * backward dx = grad_out * mask * silu'(x) with the mask regenerated from (seed [+ *seed_device], element index). */
int bdlru_silu_dropout_fwd(const void* x, void* out, int64_t n, float dropout_p, uint64_t seed,
                           const uint64_t* seed_device, int dtype, void* stream);
int bdlru_silu_dropout_bwd(const void* x, const void* grad_out, void* dx, int64_t n, float dropout_p, uint64_t seed,
                           const uint64_t* seed_device, int dtype, void* stream);

/* Column sums of a row-major [n_rows, n_cols] matrix (rows `row_stride` elements apart) -> fp32 out[n_cols]: the bias
 * gradient of the nn.Linear layers with bias (gates RecBLR.py:165; FFN RecBLR.py:213-214).  n_cols % 4 (fp32) / 8
 * (bf16) == 0. */
size_t bdlru_colsum_workspace_bytes(int64_t n_rows, int n_cols);
int bdlru_colsum(const void* x, int64_t n_rows, int n_cols, int64_t row_stride, int dtype, float* out,
                 void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Full-sort scoring with fused streaming top-k (RecBLR.py:114-122 + RecBole's scores[:,0] = -inf and
 * torch.topk).  Q [n_users, D] and E [n_rows, D] are bf16 row-major (D % 64 == 0,
 * D <= 256), accumulated in fp32 on the tcgen05 tensor cores; the [n_users, n_rows] logits are never
 * written to memory.  Row j of E has global item id id_offset + j (row-sharded tables); the item with
 * global id mask_id (pass -1 for none; RecBole masks id 0) is excluded.  Results are ordered by score
 * descending, ties broken by LOWEST item id.  1 <= k <= 32.
 *   out_scores fp32 [n_users, k], out_ids int32 [n_users, k]  (-inf / -1 when fewer than k candidates).
 * ------------------------------------------------------------------------------------------- */
/* 1 when the tcgen05 full-sort kernels are compiled into this library (else they return UNSUPPORTED). */
int bdlru_fullsort_available(void);
size_t bdlru_fullsort_topk_workspace_bytes(int64_t n_users, int64_t n_rows, int D, int k);
int bdlru_fullsort_topk(const void* Q, const void* E, int64_t n_users, int64_t n_rows, int D, int k,
                        int64_t id_offset, int64_t mask_id, float* out_scores, int32_t* out_ids,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Merge of per-shard candidate lists (the NCCL top-k merge's local step): cand_* are
 * [n_users, n_lists * k] (any order inside); keeps the k best by (score desc, id asc). */
int bdlru_topk_merge(const float* cand_scores, const int32_t* cand_ids, int64_t n_users, int n_lists, int k,
                     float* out_scores, int32_t* out_ids, void* stream);
/* Same merge reading the candidate lists IN PLACE from any regular layout: candidate (list l, user u, slot j) is element
 * l * list_stride + u * user_stride + j of cand_scores / cand_ids.  With list_stride = 2 * n_users * k, user_stride = k
 * and cand_ids = cand_scores + n_users * k it consumes the buffer an NCCL all-gather of per-rank packed
 * [2][n_users][k] (scores | ids) lists leaves behind — one collective, no permute/contiguous copies. */
int bdlru_topk_merge_strided(const float* cand_scores, const int32_t* cand_ids, int64_t n_users, int n_lists, int k,
                             int64_t list_stride, int64_t user_stride, float* out_scores, int32_t* out_ids,
                             void* stream);

/* ---------------------------------------------------------------------------------------------
 * Full-softmax cross-entropy over all item rows (RecBLR.py:99-103) without materialising the logits.
 * fwd: per user b, over this shard's rows: row_max[b], row_sumexp[b] = sum_j exp(l_bj - row_max[b]),
 *      pos_logit[b] = l_{b,pos_b} if id_offset <= pos_b < id_offset + n_rows else untouched.
 *      (single GPU: loss = mean_b(row_max + log(row_sumexp) - pos_logit); sharded: reduce first.)
 * bwd: given the global lse[b] and the upstream scale (dloss / n_users_total), recomputes the logits
 *      tile by tile and writes dQ (fp32 [n_users, D], overwritten) and dE (fp32 [n_rows, D],
 *      overwritten):  P = exp(l - lse) - onehot;  dQ = scale * P E;  dE = scale * P^T Q.
 *      scale_dev (may be NULL): DEVICE fp32 scalar multiplied into `scale` inside the kernel — the upstream gradient
 *      of the loss stays on the device and dE can be written straight into the optimizer's gradient buffer.
 *      Either of dQ / dE may be NULL (that gradient is skipped).
 * pos int64 [n_users] holds GLOBAL item ids.
 * ------------------------------------------------------------------------------------------- */
size_t bdlru_fullsort_ce_workspace_bytes(int64_t n_users, int64_t n_rows, int D);
int bdlru_fullsort_ce_fwd(const void* Q, const void* E, const int64_t* pos, int64_t n_users, int64_t n_rows,
                          int D, int64_t id_offset, float* row_max, float* row_sumexp, float* pos_logit,
                          void* workspace, size_t workspace_bytes, void* stream);
int bdlru_fullsort_ce_bwd(const void* Q, const void* E, const int64_t* pos, const float* lse, float scale,
                          const float* scale_dev, int64_t n_users, int64_t n_rows, int D, int64_t id_offset,
                          float* dQ, float* dE, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused FORWARD of the training step (one exponential pass over the logits instead of two): softmax is shift invariant,
 * so with a per-user reference m_b
 *     row_sumexp[b] = sum_j exp(l_bj - m_b),      acc[b, :] = sum_j exp(l_bj - m_b) * E[j, :]        (fp32, this shard)
 * give  lse_b = m_b + log(sum over shards of row_sumexp)  and  dL/dq_b = scale * (sum over shards of acc / s_b - E[pos_b])
 * — the dQ half of bdlru_fullsort_ce_bwd comes out of the forward, and the backward only needs the dE pass (dQ = NULL).
 * m_b must lie within ~80 of the true row maximum (fp32 / bf16 keep their relative precision over that range; beyond it
 * exp overflows to inf, which the caller detects in row_sumexp).  bdlru_fullsort_rowmax computes it on the tensor cores
 * without exponentials: exactly (tile_stride = 1) or over every tile_stride-th 96-item tile (a sampled maximum).
 * Multi-GPU: all-reduce MAX the reference before the fused pass so that every shard uses the same m_b.
 * ------------------------------------------------------------------------------------------- */
size_t bdlru_fullsort_rowmax_workspace_bytes(int64_t n_users, int64_t n_rows, int D);
int bdlru_fullsort_rowmax(const void* Q, const void* E, int64_t n_users, int64_t n_rows, int D, int tile_stride,
                          float* row_max, void* workspace, size_t workspace_bytes, void* stream);
size_t bdlru_fullsort_ce_fwd_dq_workspace_bytes(int64_t n_users, int64_t n_rows, int D);
int bdlru_fullsort_ce_fwd_dq(const void* Q, const void* E, const float* ref, int64_t n_users, int64_t n_rows, int D,
                             float* acc, float* row_sumexp, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Optimizer step of a row shard of the item table: torch.optim.Adam's update (amsgrad = False; weight_decay = L2 added
 * to the gradient) of the fp32 master rows, fused with the refresh of the bf16 compute copy (param_bf16, may be NULL):
 * one pass, 30 bytes per element instead of 34 in two kernels.  `step` is the 1-based step count (bias correction).
 * n = number of fp32 elements (multiple of 4); all arrays contiguous.
 * ------------------------------------------------------------------------------------------- */
int bdlru_table_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* param_bf16,
                          int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                          void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused core of GatedRecurrentLayer.forward for INFERENCE (RecBLR.py:182-206 without the two projections):
 *     y = silu(z) * BD-LRU(x' = silu(conv(x) + conv_b), (r | i) = gates_w x' + gates_b),   xz = (x | z)
 * as one tcgen05 kernel: reads xz [B, T, 2C] bf16 contiguous once, writes y [B, T, C] bf16; x' and the gate
 * pre-activations stay in shared memory / TMEM.  conv_w [C, 4] fp32 (NULL with conv_b NULL: no conv, x' = x), gates_w
 * [2C, C] bf16 row-major, gates_b [2C] / Lambda [C] / h0 [C] (may be NULL) fp32.  Built for C = 128 (hidden 64 x expand 2);
 * bdlru_core_fwd_supported tells.  The training path keeps the separate kernels (their backward needs x' and r|i).
 * ------------------------------------------------------------------------------------------- */
int bdlru_core_fwd_supported(int C, int dtype);
int bdlru_core_fwd(const void* xz, const float* conv_w, const float* conv_b, const void* gates_w, const float* gates_b,
                   const float* Lambda, const float* h0, void* y, int B, int T, int C, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The whole first half of RecurrentLayer.forward for INFERENCE (RecBLR.py:140-142 + 170-206) as one tcgen05 kernel:
 *     out = LayerNorm(out_w (silu(z) * BD-LRU(silu(conv(x)), gates_w . + gates_b)) + X) * ln_gamma + ln_beta,  (x | z) = in_w X
 * X [B, T, d_model] bf16 contiguous -> out [B, T, d_model] bf16; in_w [2C, d_model], gates_w [2C, C], out_w [d_model, C] bf16
 * row-major (nn.Linear layout), the rest fp32; conv_w / conv_b NULL: no conv; h0 may be NULL.  Built for d_model 64, C 128,
 * conv width 4 (bdlru_layer_fwd_supported).  Nothing between X and out touches HBM.
 * ------------------------------------------------------------------------------------------- */
int bdlru_layer_fwd_supported(int d_model, int C, int conv_width, int dtype);
int bdlru_layer_fwd(const void* x, const void* in_w, const float* conv_w, const float* conv_b, const void* gates_w,
                    const float* gates_b, const float* Lambda, const float* h0, const void* out_w, const float* ln_gamma,
                    const float* ln_beta, float eps, void* out, int B, int T, int d_model, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BDLRU_H_ */
