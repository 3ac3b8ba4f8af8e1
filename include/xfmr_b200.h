/*
 * xfmr_b200 — C ABI of the B200-native scoring-and-loss / full-catalog top-k path.
 *
 * The reference (yxtay/transformer-recommenders) is pure Python; it has no FFI layer.
 * The drop-in boundary is therefore the Python call signature of its loss modules,
 * `compute_embeds`, `compute_retrieval_metrics` and `LanceIndex.search`
 * (SURVEY.md §8b).  Those signatures are mirrored by the `xfmr_rec_b200` package,
 * which binds THIS header with ctypes and registers the entry points as
 * `torch.library` custom ops.  Each entry point cites the reference lines whose
 * arithmetic it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *     the caller owns every buffer; the library allocates nothing persistent.
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no entry
 *     point synchronises the device.
 *   - return value: 0 = ok, <0 = error (XR_E_*); `xr_last_error()` gives the
 *     message for the calling thread.
 *   - dtype codes: XR_F32 = 0, XR_BF16 = 1.  Row-major, densely packed unless a
 *     leading dimension is given.
 */
#ifndef XFMR_B200_H
#define XFMR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XR_ABI_VERSION 2

#define XR_F32 0
#define XR_BF16 1

#define XR_OK 0
#define XR_E_INVALID (-1)     /* bad argument (shape, dtype, alignment, null pointer)      */
#define XR_E_CUDA (-2)        /* a CUDA runtime / driver call failed                        */
#define XR_E_UNSUPPORTED (-3) /* valid request this build cannot serve (e.g. not sm_100)    */

/* loss kinds; order = LOSS_CLASSES, xfmr_rec/losses.py:546-554 */
#define XR_LOSS_ALIGNMENT 0             /* losses.py:408-426 */
#define XR_LOSS_ALIGNMENT_CONTRASTIVE 1 /* losses.py:429-447 (CCL)  */
#define XR_LOSS_CONTRASTIVE 2           /* losses.py:450-469 */
#define XR_LOSS_INFONCE 3               /* losses.py:472-488 (SSM)  */
#define XR_LOSS_NCE 4                   /* losses.py:491-511 */
#define XR_LOSS_PAIRWISE_HINGE 5        /* losses.py:514-527 */
#define XR_LOSS_PAIRWISE_LOGISTIC 6     /* losses.py:530-543 (BPR at margin 0) */
#define XR_NUM_LOSSES 7

/* target_position, losses.py:26, 240-253 */
#define XR_TARGET_FIRST 0
#define XR_TARGET_DIAGONAL 1
#define XR_TARGET_EXPLICIT 2
#define XR_TARGET_LAST 3 /* internal layout of xr_logits_pool */

/* LossConfig, xfmr_rec/losses.py:11-30 (target_position carried separately) */
typedef struct xr_loss_config {
  int32_t mask_false_negatives; /* losses.py:27, 283-292 */
  int32_t num_hard_negatives;   /* losses.py:28, 311-330 */
  float scale;                  /* losses.py:29, 486     */
  float margin;                 /* losses.py:30, 370, 525, 541 */
  int32_t logits_bf16;          /* 1: round dot logits to bf16 before masking, as Lightning's
                                   bf16-mixed autocast does to losses.py:195 (trainer.py:450) */
} xr_loss_config;

/* number of float64 slots in the statistics block written by xr_rowloss / fused kernels */
#define XR_STATS_SLOTS 16

const char* xr_last_error(void);
int xr_abi_version(void);
/* sm count, compute capability and whether the tcgen05 kernels can run on the current device */
int xr_device_info(int* sm_count, int* cc_major, int* cc_minor, int* has_tcgen05);
/* Leave `n_reserved` SMs free: the persistent kernels (one CTA per SM) size their grids with the physical
 * SM count minus this, so that kernels of OTHER streams -- the NCCL all-reduce of the data-parallel
 * trainer's encoder gradients -- find a place to run under them instead of behind them.  Returns the
 * previous value; n_reserved < 0 only reads it.  Process-wide; set it before capturing CUDA graphs.      */
int xr_reserve_sms(int n_reserved);

/* ---- family 1: embedding row gathers ------------------------------------------------------
 * nn.Embedding.forward — models.py:336-338 (history), :400 (positives), :406 (negatives).
 * out[i,:] = table[ idx[ sel ? sel[i] : i ], : ].  Bit-exact copy when out_dtype ==
 * table_dtype; XR_F32 -> XR_BF16 rounds to nearest even (torch's .bfloat16()).
 * `sel` (nullable) is the second-level index produced by xr_compact_positions, which fuses
 * the boolean-mask compaction of models.py:398/404 into the gather.
 * `err_flag` (nullable, device int32): set to 1 if any index is outside [0, n_rows).          */
int xr_gather_rows(const void* table, int64_t n_rows, int64_t dim, int table_dtype,
                   const int64_t* idx, const int64_t* sel, int64_t n_out, void* out,
                   int out_dtype, int32_t* err_flag, void* stream);

/* dst[sel[i],:] = src[i,:] (dst pre-zeroed by the caller, sel unique) — the autograd transpose
 * of the query compaction `token_embeddings[attention_mask][pos_mask]` (models.py:392, 415).  */
int xr_scatter_rows(const float* src, int64_t n_src, int64_t dim, const int64_t* sel,
                    float* dst, int64_t n_dst_rows, void* stream);

/* dst[p,:] = inv_pos[p] >= 0 ? cast(src[inv_pos[p],:] * *scale) : 0 — zero fill + scatter + the
 * grad_output scale (device scalar, nullable = 1) + cast in one pass; dst is (n_dst_rows, dim).  */
int xr_scatter_scaled(const float* src, const int64_t* inv_pos, const float* scale,
                      int64_t n_dst_rows, int64_t dim, void* dst, int dst_dtype, void* stream);

/* rownz[r] = any(table[r,:] != 0) — lets the attention mask of models.py:343 be derived from
 * the indices: mask = rownz[idx].                                                             */
int xr_row_nonzero(const void* table, int64_t n_rows, int64_t dim, int dtype, uint8_t* rownz,
                   void* stream);

/* models.py:343, 390, 398, 404, 413-416 on the index tensors of one SeqBatch (flattened B*L):
 *   attn[p]      = rownz[history_idx[p]]                      (attention_mask, :343/:390)
 *   sel_attn[..] = positions p with attn[p], ascending        (neg / pos row order, :398/:404)
 *   sel_pos[..]  = positions p with attn[p] && pos_idx[p]!=0  (query / candidate rows, :413-416)
 *   pos_mask[a]  = pos_idx[sel_attn[a]] != 0                   (positive_mask, :413)
 *   inv_pos[p]   = row of position p in sel_pos, or -1      (nullable; inverse map for the backward)
 * counts[0] = M_a, counts[1] = M (device int64[2]).  rownz == NULL uses idx != 0 instead.
 * workspace >= xr_compact_workspace_bytes(n_pos).                                             */
size_t xr_compact_workspace_bytes(int64_t n_pos);
int xr_compact_positions(const int64_t* history_idx, const int64_t* pos_idx,
                         const uint8_t* rownz, int64_t n_table_rows, int64_t n_pos,
                         uint8_t* attn, int64_t* sel_attn, int64_t* sel_pos, uint8_t* pos_mask,
                         int64_t* inv_pos, int64_t* counts, void* workspace, void* stream);

/* y[r,:] = x[r,:] / max(||x[r,:]||, eps); inv_norm[r] = 1/max(||x[r,:]||, eps);
 * the two normalisations inside torch's cosine_similarity, losses.py:206-208 (eps 1e-8) and
 * the catalog/query normalisation of the cosine index metric, index.py:47.
 * y may be null (norms only); y dtype may differ from x dtype.                               */
int xr_normalize_rows(const void* x, int64_t n_rows, int64_t dim, int x_dtype, float eps,
                      void* y, int y_dtype, float* inv_norm, void* stream);

/* ---- family 2a: logits, materialised (fp32-exact path and API-compat paths) ---------------
 * compute_logits / cosine_similarity_logits, losses.py:179-209.                               */

/* shared-pool candidates (models.py:408-410 without the O(M^2 D) copy), fp32 accumulate:
 *   logits[i,j] = q_i . neg_j  for j < Cn ;  logits[i,Cn] = q_i . pos_i
 * The positive sits in the LAST column (XR_TARGET_LAST) so both GEMM operands stay 16-byte
 * aligned; masking, mining and every loss are invariant to where the target column is.
 * logits is (M, ld) fp32 with ld >= Cn+1 (ld % 4 == 0 keeps the vector path).                 */
int xr_logits_pool(const void* q, const void* pos, const void* neg, int64_t m, int64_t cn,
                   int64_t dim, int dtype, float* logits, int64_t ld, void* stream);

/* genuine dense candidates (M,C,D) — the reference's bmm, losses.py:195 (batched GEMV).
 * cosine (losses.py:206-208): pass q_inv_norm (from xr_normalize_rows) and a (M,C) fp32
 * cand_inv_norm_out buffer, filled in the same pass over the candidates; both NULL for dot.  */
int xr_logits_dense(const void* q, const void* cand, int64_t m, int64_t c, int64_t dim,
                    int dtype, const float* q_inv_norm, float* cand_inv_norm_out, float eps,
                    float* logits, int64_t ld, void* stream);

/* per-row sampled negatives: candidates of row i are table[cand_idx[i,0..c)] (col 0 = positive)
 * — BASELINE config 3; fused gather + dot, no (M,C,D) tensor.                                 */
int xr_logits_sampled(const void* q, const void* table, int64_t n_rows, const int64_t* cand_idx,
                      int64_t m, int64_t c, int64_t dim, int dtype, const float* table_inv_norm,
                      const float* q_inv_norm, float* logits, int64_t ld, void* stream);

/* ---- the EmbedLoss pipeline on logits: losses.py:211-330 + the seven loss() bodies ---------
 * check_target (:233-261) -> mask_false_negatives (:283-292) -> mine_hard_negatives (:311-330)
 * -> loss (:420-543) and LogitsStatistics (:383-405), one pass per row.
 *   losses_out  : float64[XR_NUM_LOSSES] sums over rows for every loss whose bit is set in
 *                 `loss_mask` (bit k = loss kind k); logits are used as given, so cosine and
 *                 dot families are evaluated in separate calls.
 *   stats_out   : float64[XR_STATS_SLOTS] (nullable): [0] sum_i n_valid_i/(num_neg+1e-9),
 *                 [1] rows, pos {[2] sum,[3] sumsq,[4] min,[5] max}, neg {[6] count,[7] sum,
 *                 [8] sumsq,[9] min,[10] max}, [11] num_neg used for the density.
 *   dlogits     : (M, ld) fp32 (nullable): dL/dlogits of `grad_kind` (-1 = none), times
 *                 grad_scale.
 *   err_flag    : nullable device int32, set when an explicit target is out of range (the
 *                 reference's gather would raise).
 *   workspace   : >= xr_rowloss_workspace_bytes(m, c, cfg->num_hard_negatives) bytes.
 * Hard-negative ties are resolved toward the lower candidate index (torch.topk(sorted=False)
 * leaves it implementation-defined, losses.py:318-322).                                       */
size_t xr_rowloss_workspace_bytes(int64_t m, int64_t c, int num_hard_negatives);
int xr_rowloss(const float* logits, int64_t m, int64_t c, int64_t ld, int target_mode,
               const int64_t* target, const xr_loss_config* cfg, uint32_t loss_mask,
               int grad_kind, float grad_scale, float* dlogits, double* losses_out,
               double* stats_out, int32_t* err_flag, void* workspace, void* stream);

/* ---- sampled candidates in ONE pass (BASELINE config 3) --------------------------------------
 * xr_logits_sampled + xr_rowloss (target = column 0) + xr_dq_sampled for one query row per
 * thread block, the row's C logits and dL/dlogits kept in shared memory: the (M, C) logits and
 * their gradient never reach HBM and the step is one launch plus the fixed-order reduction.
 * D = 384, 1 <= c <= 8192.  cosine iff table_inv_norm / q_inv_norm are given (both or neither).
 * dq: (M, D) fp32, written when grad_kind >= 0.  losses_out / stats_out as xr_rowloss.
 * workspace: >= xr_sampled_step_workspace_bytes(m) bytes.                                      */
size_t xr_sampled_step_workspace_bytes(int64_t m);
int xr_sampled_step(const void* q, const void* table, int64_t n_rows, const int64_t* cand_idx,
                    int64_t m, int64_t c, int64_t dim, int dtype, const float* table_inv_norm,
                    const float* q_inv_norm, const xr_loss_config* cfg, int grad_kind,
                    float grad_scale, float* dq, double* losses_out, double* stats_out,
                    void* workspace, void* stream);

/* ---- family 2b: dL/dquery from dL/dlogits ---------------------------------------------------
 * autograd of losses.py:195 / :206-208 w.r.t. query_embed (item table frozen, models.py:251).
 * xr_dq_pool, cosine != 0: q/pos/neg are the NORMALISED rows plus q_inv_norm; it applies
 * dq = inv_norm * (g - (g.qhat) qhat).  The dense / sampled variants take the RAW q plus the
 * inverse norms and normalise on the fly.                                                     */
int xr_dq_pool(const float* dlogits, int64_t ld, const void* q, const void* pos, const void* neg,
               int64_t m, int64_t cn, int64_t dim, int dtype, int cosine,
               const float* q_inv_norm, float* dq, void* stream);
int xr_dq_dense(const float* dlogits, int64_t ld, const void* q, const void* cand, int64_t m,
                int64_t c, int64_t dim, int dtype, int cosine, const float* q_inv_norm,
                const float* cand_inv_norm, float* dq, void* stream);
/* dL/d candidate_embed for a dense (M, C, D) candidate tensor (losses.py:128-155 is differentiable in both
 * arguments; the trainer never needs it: models.py:251-253 freezes the table).  dot: w * q_i; cosine (pass the
 * forward's logits and both inverse norms): w * (q^ - cos * c^) / |c|.  dcand: (M, C, D) fp32.             */
int xr_dcand_dense(const float* dlogits, const float* logits, int64_t ld, const void* q, const void* cand,
                   int64_t m, int64_t c, int64_t dim, int dtype, int cosine, const float* q_inv_norm,
                   const float* cand_inv_norm, float* dcand, void* stream);
int xr_dq_sampled(const float* dlogits, int64_t ld, const void* q, const void* table,
                  int64_t n_rows, const int64_t* cand_idx, int64_t m, int64_t c, int64_t dim,
                  int dtype, const float* table_inv_norm, const float* q_inv_norm, float* dq,
                  void* stream);

/* ---- the monitoring half of compute_losses inside the sync-free step -------------------------
 * trainer.py:250-263 logs LogitsStatistics and all seven losses on every training step.  Call
 * right after xr_pool_step on the SAME stream with the SAME workspace (sized by
 * xr_pool_step_monitor_workspace_bytes, which xr_pool_step accepts too): both all-losses passes
 * (xr_fused_pool_all, dot and cosine family) run on the operands the step gathered, the shape
 * stays on the device, nothing is copied to the host -- CUDA-graph capturable with the step.
 *   losses_dot / losses_cos : float64[XR_NUM_LOSSES] (entries of the other family meaningless)
 *   stats_out               : float64[XR_STATS_SLOTS] of the dot logits (losses.py:383-405)
 * cfg->logits_bf16 applies to the dot family; cosine logits stay fp32.                          */
size_t xr_pool_step_monitor_workspace_bytes(int64_t n_pos, int64_t dim);
int xr_pool_step_monitor(int64_t n_pos, int64_t dim, const xr_loss_config* cfg, double* losses_dot,
                         double* losses_cos, double* stats_out, void* workspace,
                         size_t workspace_bytes, void* stream);
/* The compute phase of the step WITH the monitoring folded into the train kernel (trainer.py:250-263 +
 * 288-300 from ONE tensor-core pass): train loss, dL/dtok, losses_dot[7], losses_cos[7], stats[16].  Dot-family
 * train losses (InfoNCE, NCE, pairwise hinge / logistic), scale > 0, no hard-negative mining; call after xr_pool_step_ingest on the same stream with a
 * workspace of xr_pool_step_monitor_workspace_bytes.  The dot family equals xr_pool_step_monitor bit for bit;
 * the cosine family is evaluated as (score / |q|) / |n| with fp32 inverse norms.                          */
int xr_pool_step_compute_mon(int64_t n_pos, int64_t dim, int loss_kind, const xr_loss_config* cfg,
                             float grad_scale, void* dtok, int dtok_dtype, double* loss_out, double* losses_dot,
                             double* losses_cos, double* stats_out, void* workspace, size_t workspace_bytes,
                             void* stream);
/* The same one-pass evaluation for exact shapes (the module path of trainer.py:213-264, compute_losses):
 * train loss of a dot-family kind (loss_out[0], fp32 copy in the low half of loss_out[1]) and dq (m, dim) fp32 (nullable),
 * losses_dot[7], losses_cos[7], stats[16] from ONE tensor-core pass.  bf16 operands, dim 384.            */
size_t xr_fused_pool_loss_mon_workspace_bytes(int64_t m, int64_t cn, int64_t dim);
int xr_fused_pool_loss_mon(const void* q, const void* pos, const void* neg, int64_t m, int64_t cn, int64_t dim,
                           int loss_kind, const xr_loss_config* cfg, float grad_scale, float* dq, double* loss_out,
                           double* losses_dot, double* losses_cos, double* stats_out, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ---- SeqBatch construction (the step right before the path) ---------------
 * SeqDataset.__getitem__ + collate (data.py:669-805) for a whole batch in one launch:
 * sample_sequence (:669-689), sample_positives (:691-721), sample_negatives (:723-747),
 * zero right-padding (:787-805).  Inputs are the per-user event histories in CSR form, resident
 * on the device:
 *   hist_off[H+1], items[total] (1-based item idx), labels[total], pos_prefix[total] (inclusive
 *   count of positive labels inside each history), uniq_off[H+1] / uniq_items (ascending unique
 *   items of each history), row_hist[R] (dataset row -> history, data.py:618-636; NULL = identity),
 *   rows[n_batch] (the dataset rows of this batch).
 * Outputs: history_out / pos_out / neg_out int64 (n_batch, max_seq_length), 0-padded;
 *   seq_len_out int32[n_batch] (nullable).
 * Randomness: Philox4x32-10 keyed by (seed, step, dataset row): deterministic, independent of
 * the batch composition; oracle/xfmr_oracle.py:seq_sample_batch reproduces it bit for bit.     */
int xr_seq_sample_batch(const int64_t* hist_off, const int64_t* items, const uint8_t* labels,
                        const int32_t* pos_prefix, const int64_t* uniq_off,
                        const int64_t* uniq_items, const int64_t* row_hist, const int64_t* rows,
                        int64_t n_batch, int64_t n_items, int max_seq_length, int pos_lookahead,
                        uint64_t seed, uint64_t step, int64_t* history_out, int64_t* pos_out,
                        int64_t* neg_out, int32_t* seq_len_out, void* stream);

/* ---- family 2c: fused contraction + loss + gradient (tcgen05/TMEM, bf16) -------------------
 * One pass over the shared negative pool: S = Q.Neg^T on the 5th-gen tensor cores, the loss
 * epilogue on the TMEM accumulator, dQ += W.Neg as a second UMMA — the M x C logits never
 * reach HBM.  Replaces models.py:408-416 + losses.py:150-155 for InfoNCE (:479-488),
 * PairwiseLogistic / PairwiseHinge (:520-543), NCE (:498-511) and, on pre-normalised inputs,
 * Contrastive / AlignmentContrastive (:338-372, :436-447).  num_hard_negatives must be 0.
 *   q, pos : (M, D) bf16;  neg : (Cn, D) bf16;  D = 384 in this build.
 *   dq     : (M, D) fp32 (nullable => forward only), times grad_scale.
 *   loss_out : float64[2]: [0] sum over rows; the first 4 bytes of [1] receive its fp32 copy.
 *   row_loss (nullable): float32[M].
 *   q_inv_norm (cosine kinds only): float32[M].
 *   workspace: >= xr_fused_pool_workspace_bytes(m, cn, dim) bytes.                            */
int xr_fused_available(void); /* bit 0: fused loss kernel, bit 1: fused score+top-k kernel */
/* measurement hook: record a CUDA event pair around every main fused-kernel launch (ring of
 * 512); xr_fused_profile_read copies the durations in ms to HOST memory and resets the ring. */
int xr_fused_profile(int enable);
int xr_fused_profile_read(float* ms_out_host, int max_n);
/* profiling aid: per-barrier wait cycles of the last fused call (16 host counters, see the .cu) */
int xr_fused_wait_stats(int enable, unsigned long long* out16_host);
int xr_fused_timeline(long long* out512_host); /* profiling aid: per-tile timestamps of CTA 0 */
size_t xr_fused_pool_workspace_bytes(int64_t m, int64_t cn, int64_t dim);
int xr_fused_pool_loss(const void* q, const void* pos, const void* neg, int64_t m, int64_t cn,
                       int64_t dim, int loss_kind, const xr_loss_config* cfg,
                       const float* q_inv_norm, float grad_scale, float* dq, double* loss_out,
                       float* row_loss, void* workspace, size_t workspace_bytes, void* stream);

/* ---- family 2d: every loss of one logit family + LogitsStatistics in one tensor-core pass ----
 * RecommenderLightningModule.compute_losses (trainer.py:250-263) evaluates LogitsStatistics
 * (losses.py:383-405) and all seven losses on every training step: eight logit computations in
 * the reference.  This call produces, forward only and without materialising the logits,
 *   cosine == 0 (q/pos/neg as given, dot logits, losses.py:195): InfoNCE, NCE, PairwiseHinge,
 *                PairwiseLogistic and the statistics block;
 *   cosine != 0 (q/pos/neg pre-normalised rows, losses.py:206-208): Alignment, Contrastive,
 *                AlignmentContrastive (statistics: density / counts / positives only --
 *                LogitsStatistics is defined on the dot logits, losses.py:383-386),
 * in the same float64[XR_NUM_LOSSES] / float64[XR_STATS_SLOTS] layout as xr_rowloss; entries
 * of the other family are meaningless (as with xr_rowloss).  num_hard_negatives must be 0.   */
size_t xr_fused_pool_all_workspace_bytes(int64_t m, int64_t cn, int64_t dim);
int xr_fused_pool_all(const void* q, const void* pos, const void* neg, int64_t m, int64_t cn,
                      int64_t dim, int cosine, const xr_loss_config* cfg, double* losses_out,
                      double* stats_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- the whole scoring-and-loss step, sync-free ----------------------------------------------
 * RecommenderModel.compute_embeds (models.py:388-416) + EmbedLoss.forward (losses.py:128-155)
 * + the backward to the encoder output (autograd of models.py:392, 415), for ONE SeqBatch of
 * n_pos = B*L positions, in one stream-ordered call with no device->host copy: the row counts
 * M_a / M stay on the device (`counts`, int64[2], also an output), the tensor-core kernel is
 * planned there, and every buffer is sized by n_pos — so the launch sequence is static and the
 * call can be captured into a CUDA graph.  The same kernels, in the same order, as
 * xr_compact_positions + xr_gather_rows x3 + xr_fused_pool_loss + xr_scatter_scaled: results
 * are bit-identical to that sequence (tests/test_gpu_step.py).
 *   history_idx / pos_idx / neg_idx: int64[n_pos] (SeqBatch index tensors, flattened B*L)
 *   tok:       (n_pos, dim) encoder output, XR_F32 or XR_BF16 (rounded to bf16 operands)
 *   table_bf16:(n_table_rows, dim) bf16 item table, row 0 = padding;  rownz: xr_row_nonzero of
 *              the fp32 master table (nullable: idx != 0)
 *   loss_kind: InfoNCE / NCE / PairwiseHinge / PairwiseLogistic (dot-product kinds)
 *   dtok:      (n_pos, dim) dL/d tok scaled by grad_scale, zero rows for unselected positions;
 *              nullable = forward only.  loss_out: double[2] ([0] f64 sum, [1] low word = f32).
 *   err_flag:  nullable device int32, set to 1 on an out-of-range item index.
 *   workspace: >= xr_pool_step_workspace_bytes(n_pos, dim) bytes, 256-byte aligned.             */
size_t xr_pool_step_workspace_bytes(int64_t n_pos, int64_t dim);
int xr_pool_step(const int64_t* history_idx, const int64_t* pos_idx, const int64_t* neg_idx,
                 int64_t n_pos, const void* tok, int tok_dtype, const void* table_bf16,
                 const uint8_t* rownz, int64_t n_table_rows, int64_t dim, int loss_kind,
                 const xr_loss_config* cfg, float grad_scale, void* dtok, int dtok_dtype,
                 double* loss_out, int64_t* counts, int32_t* err_flag, void* workspace,
                 size_t workspace_bytes, void* stream);

/* The same step in two phases for callers that pipeline batches: INGEST (index compaction, device
 * plan, the three gathers: everything that reads the batch's inputs) and COMPUTE (diagonal pass,
 * fused loss forward/backward, finalize, row sum).  xr_pool_step == ingest + compute on one stream.
 * With two alternating workspaces the ingest of batch i+1 runs on a second stream under the compute
 * of batch i.  `tok` may point to pinned HOST memory (UVA): the gather then pulls only the selected
 * M rows over PCIe instead of a whole (B, L, D) copy.                                          */
int xr_pool_step_ingest(const int64_t* history_idx, const int64_t* pos_idx, const int64_t* neg_idx,
                        int64_t n_pos, const void* tok, int tok_dtype, const void* table_bf16,
                        const uint8_t* rownz, int64_t n_table_rows, int64_t dim, int64_t* counts,
                        int32_t* err_flag, void* workspace, size_t workspace_bytes, void* stream);
int xr_pool_step_compute(int64_t n_pos, int64_t dim, int loss_kind, const xr_loss_config* cfg,
                         float grad_scale, void* dtok, int dtok_dtype, double* loss_out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---- family 3: top-k -------------------------------------------------------------------------
 * Exact top-k of each row of a materialised (U,N) fp32 score matrix under the total order
 * (score descending, column ascending) — the result the reference's ANN search
 * (index.py:244-251, limit(top_k)) approximates; ties -> lower item id.
 * col_offset is added to the reported indices (catalog shard base).  NaN ranks first
 * (torch.sort convention).  out_scores (U,k) fp32, out_idx (U,k) int64; if N < k the tail is
 * filled with -inf / -1.                                                                      */
size_t xr_topk_workspace_bytes(int64_t u, int64_t n, int64_t k);
int xr_topk(const float* scores, int64_t u, int64_t n, int64_t ld, int64_t k, int64_t col_offset,
            float* out_scores, int64_t* out_idx, void* workspace, size_t workspace_bytes,
            void* stream);
/* merge G per-shard results laid out (U, G*k) (scores + global ids), same total order; the
 * local step after the NCCL all-gather of SURVEY §8e.                                         */
size_t xr_topk_merge_workspace_bytes(int64_t u, int64_t k);
int xr_topk_merge(const float* scores, const int64_t* ids, int64_t u, int64_t gk, int64_t k,
                  float* out_scores, int64_t* out_idx, void* workspace, void* stream);

/* Sharded retrieval exchange as ONE kernel over peer memory: all-gather + merge without NCCL.
 * peer_scores_host / peer_ids_host: HOST arrays of n_peers DEVICE pointers, entry g addressing
 * rank g's (U, k) fp32 scores / int64 global ids (ids < 2^32, -1 = none) through a peer mapping
 * (CUDA IPC / symmetric memory; entry `rank` is the caller's own buffer).  The caller orders the
 * kernel after a cross-GPU barrier that follows every rank's writes.  Same total order as
 * xr_topk_merge (score desc, id asc): the result equals the single-GPU search bit for bit.
 * workspace: >= xr_topk_merge_workspace_bytes(u, k).                                          */
int xr_topk_merge_peers(const void* const* peer_scores_host, const void* const* peer_ids_host,
                        int n_peers, int64_t u, int64_t k, float* out_scores, int64_t* out_idx,
                        void* workspace, void* stream);

/* scores[u,n] = q_u . cat_n (* q_inv_norm[u] * cat_inv_norm[n] when given), fp32 accumulate;
 * exclusion: for every (u, e) pair in the CSR lists excl_offsets/excl_ids (global ids), the
 * score of column e - col_offset is set to -inf (the prefilter of index.py:239-247).          */
int xr_scores(const void* q, int64_t u, const void* catalog, int64_t n, int64_t dim, int dtype,
              const float* q_inv_norm, const float* cat_inv_norm, float* scores, int64_t ld,
              void* stream);
int xr_mask_excluded(float* scores, int64_t u, int64_t n, int64_t ld, int64_t col_offset,
                     const int64_t* excl_offsets, const int64_t* excl_ids, void* stream);

/* Retrieval scoring on the tensor cores without materialising the (U, N) score matrix:
 *   gmax[u, g] = max over the 16 catalog rows of group g of q_u . cat_c   (fp32; -inf past n)
 * q (U, 384) and catalog (N, 384) bf16 (pre-normalised rows for the cosine metric, index.py:47).
 * Tiles of T catalog rows (T = 128 for U > 128 — the cta_group::2 kernel — else 64); with
 * tile_stride = s only every s-th tile is scored (a SAMPLE of the catalog).  Storage column
 * (T/16) t + g of gmax (U, ld) holds rows [T t s + 16 g, +16): for s = 1 column c = rows [16 c, +16), so
 * xr_topk's (maximum desc, column asc) order is (maximum desc, first row asc).
 * The k-th largest group maximum of a row — of the whole catalog or of any sample of it — is a lower
 * bound of the row's k-th largest score (index.py:244-254, `.limit(top_k)`, exact).
 * ld: even, >= xr_score_groupmax_ld(u, n, tile_stride); every one of those columns is written.       */
int64_t xr_score_groupmax_ld(int64_t u, int64_t n, int64_t tile_stride);
int xr_score_groupmax(const void* q, int64_t u, const void* catalog, int64_t n, int64_t dim,
                      int64_t tile_stride, float* gmax, int64_t ld, void* stream);
/* Threshold filter in the scoring epilogue: every (score, local row) with
 * q_u . cat_c >= thresh[u * thresh_stride] is kept.  Storage per query: n_sub sub-buckets of cap_b slots
 * (bucket_scores / bucket_rows, (U, n_sub, cap_b)), one per (catalog split of the scoring plan, column
 * group) — each filled by ONE lane of the scoring kernel from a register counter, no atomics — plus an
 * overflow list of ovf_cap slots for sub-buckets that run full.  bucket_count[u][s] = survivors
 * sub-bucket s saw (> cap_b: the excess is in the overflow list); ovf_count[u] (zeroed by the caller)
 * keeps counting past ovf_cap: ovf_count > ovf_cap means survivors were lost.  n_sub must be the value
 * xr_score_filter_layout returns for (u, n); cap_b is the caller's choice (the layout call suggests
 * ~4x the expected fill).  Nothing of size U x N reaches HBM.                                          */
int xr_score_filter_layout(int64_t u, int64_t n, int64_t expected_survivors, int64_t* n_sub,
                           int64_t* cap_b);
int xr_score_filter(const void* q, int64_t u, const void* catalog, int64_t n, int64_t dim,
                    const float* thresh, int64_t thresh_stride, float* bucket_scores,
                    int32_t* bucket_rows, int32_t* bucket_count, int64_t n_sub, int64_t cap_b,
                    float* ovf_scores, int32_t* ovf_rows, int32_t* ovf_count, int64_t ovf_cap,
                    void* stream);
/* Survivors -> the exact top-k (block per query): the k_sel best survivors under
 * (score desc, row asc) are kept (one radix select in shared memory), rows in query u's CSR exclusion
 * list (GLOBAL ids; nullable) are dropped (the prefilter of index.py:239-247), the rest is ranked by
 * (score desc, global id asc).  The scores reported are the tensor-core scores the filter stored: the
 * threshold, the selection and the ranking share one arithmetic.  out_scores (U, k) fp32 / out_idx (U, k)
 * int64 = local row + row_offset (-inf / -1 where fewer than k remain).  thresh = the thresholds the
 * survivors were filtered with.  flags[0] |= 1 if survivors were lost (ovf_count > ovf_cap), |= 2 if a
 * query has more than max_excl excluded ids, |= 4 if fewer than k non-excluded rows survived a finite
 * threshold that kept at most k_sel rows: the result is then not guaranteed exact and the caller must
 * take another path.  k + max_excl <= k_sel <= 1024; n = catalog rows (global ids must be < 2^32).      */
int xr_filter_finalize(int64_t u, int64_t n, const float* bucket_scores, const int32_t* bucket_rows,
                       const int32_t* bucket_count, int64_t n_sub, int64_t cap_b,
                       const float* ovf_scores, const int32_t* ovf_rows, const int32_t* ovf_count,
                       int64_t ovf_cap, const float* thresh, int64_t thresh_stride, int64_t k_sel,
                       int64_t k, int64_t row_offset, const int64_t* excl_offsets,
                       const int64_t* excl_ids, int64_t max_excl, float* out_scores, int64_t* out_idx,
                       int32_t* flags, void* stream);
/* out[r] = the kth largest of x[r, 0..n) (fp32, NaN ranks first), -inf when n < kth: the thresholds of the
 * filter from the sample's group maxima, one block per row.                                            */
int xr_kth_largest(const float* x, int64_t u, int64_t n, int64_t ld, int64_t kth, float* out, void* stream);
/* scores[u, j] = -inf where ids[u, j] lies outside [id_lo, id_hi) or in row u's CSR exclusion
 * list (nullable) — the prefilter of index.py:239-247 applied to re-scored candidate lists.     */
int xr_mask_excluded_ids(float* scores, const int64_t* ids, int64_t u, int64_t c, int64_t ld,
                         int64_t id_lo, int64_t id_hi, const int64_t* excl_offsets,
                         const int64_t* excl_ids, void* stream);
/* group ids (u, kg) chosen by xr_topk over the group maxima of xr_score_groupmax (tile_stride 1) ->
 * the 16 catalog rows of each group: cols (u, kg*16) local rows for the re-score gather (-1 where
 * there is no such row), ids (u, kg*16) global row ids = local + row_offset, -1 where none.     */
int xr_groups_to_rows(const int64_t* group_ids, int64_t u, int64_t kg, int64_t n, int64_t row_offset,
                      int64_t* cols, int64_t* ids, void* stream);

/* The whole local search of one catalog shard as ONE call (bf16, dim 384) — LanceIndex.search /
 * FaissIndex.search, index.py:214-255 / 439-474, exact and batched over U queries:
 *   sample group maxima (xr_score_groupmax, tile_stride s) -> the (k + 28)-th largest per query =
 *   threshold (xr_kth_largest) -> xr_score_filter over the whole shard -> xr_filter_finalize with
 *   k_sel = k + max_excl.
 * Same result as xr_scores + xr_mask_excluded + xr_topk unless flags[0] != 0 afterwards (see
 * xr_filter_finalize; flags is NOT cleared by the call, so one word can watch many searches).
 * q / catalog rows pre-normalised for the cosine metric.  excl_*: nullable CSR of GLOBAL ids, at most
 * max_excl per query.  workspace: >= xr_score_topk_workspace_bytes(u, n, k, max_excl), 256-byte aligned. */
size_t xr_score_topk_workspace_bytes(int64_t u, int64_t n, int64_t k, int64_t max_excl);
int xr_score_topk(const void* q, int64_t u, const void* catalog, int64_t n, int64_t dim, int64_t k,
                  int64_t row_offset, const int64_t* excl_offsets, const int64_t* excl_ids,
                  int64_t max_excl, float* out_scores, int64_t* out_idx, int32_t* flags,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ---- retrieval metrics -------------------------------------------------------------------------
 * compute_retrieval_metrics, metrics.py:62-79 (+ torchmetrics 1.9.0 functional definitions),
 * batched: rec (U,k) int64 ranked ids (-1 = the "" padding of metrics.py:65-68), targets in CSR.
 * out (U,7) fp32 in METRIC_FNS order (metrics.py:6-14); valid[u] = 0 where the user has no
 * targets (the reference returns {} there, metrics.py:62-63).                                 */
int xr_retrieval_metrics(const int64_t* rec, int64_t u, int64_t k, const int64_t* tgt_offsets,
                         const int64_t* tgt_ids, int64_t top_k, float* out, uint8_t* valid,
                         void* stream);

/* ---- sequence encoder -------------------------------------------------------------
 * The reference's encoder is a HuggingFace BertModel(is_decoder=True) fed `inputs_embeds` from the frozen
 * item table (models.py:51-102, 306-345).  These are the fused kernels around its linear layers (which
 * are plain GEMMs and stay with cuBLAS): hidden size 384, head_dim 32, sequences of at most 384 positions.
 * Activations `dtype`: XR_F32 or XR_BF16; LayerNorm statistics, softmax and every reduction are fp32.
 *
 * embed_ln: out[b,l,:] = LayerNorm(table[idx[b,l]] + pos_emb[l] + type_emb[0:H]) (BertEmbeddings with
 * inputs_embeds; the history gather of models.py:336-338 fused in), stats (B*L, 2) = {mean, rstd},
 * mask[b,l] = any(table[idx[b,l]] != 0) (the attention mask of models.py:343).                           */
/* Dropout (HF BertConfig: hidden_dropout_prob / attention_probs_dropout_prob, 0.1 by default, training mode):
 * every entry point below that has a dropout site takes `rng` = device pointer to {int64 seed, int64 step
 * counter} (NULL = no dropout), the probability and a site id.  The keep mask is a pure function of (seed,
 * counter, site, element) through Philox4x32-10: the backward entry recomputes it from the SAME rng values,
 * nothing is stored, and a CUDA-graph replay draws a fresh mask when the caller bumps the counter on the
 * device.  Hidden-state sites use 16-bit draws (p_eff = round(65536 p) / 65536), the attention-probability
 * site 8-bit draws (p_eff = round(256 p) / 256); kept values are scaled by 1 / (1 - p_eff).              */
size_t xr_enc_ln_workspace_bytes(int64_t n_tok);
int xr_enc_embed_ln_fwd(const float* table, int64_t n_table_rows, const int64_t* idx, const float* pos_emb,
                        const float* type_emb, const float* gamma, const float* beta, int64_t batch,
                        int64_t seq_len, int64_t dim, float eps, float* out, void* out_bf16 /* nullable */,
                        float* stats, uint8_t* mask, int32_t* err_flag, const int64_t* rng, float drop_p, int site,
                        void* stream);
/* gradients of the position embeddings (seq_len, H), token-type row 0 (H), LayerNorm weight / bias; the item
 * table is frozen (models.py:251-253).  The upstream gradient is dout (fp32) + dout_bf16 (the gradient of the
 * bf16 copy), either may be NULL.  workspace: xr_enc_ln_workspace_bytes(batch * seq_len).                  */
int xr_enc_embed_ln_bwd(const float* table, int64_t n_table_rows, const int64_t* idx, const float* pos_emb,
                        const float* type_emb, const float* gamma, const float* stats, const float* dout,
                        const void* dout_bf16, int64_t batch, int64_t seq_len, int64_t dim, float* dpos,
                        float* dtype0, float* dgamma, float* dbeta, void* workspace, const int64_t* rng, float drop_p,
                        int site, void* stream);
/* out = LayerNorm(y + bias + residual) (BertSelfOutput / BertOutput; y = the dense layer's product, its bias
 * added here; bias may be NULL when y already holds it).  out_bf16 (nullable): the same rows rounded to bf16,
 * the next GEMM's input.  Backward: dresidual (fp32) and dy (y's dtype) both receive the LayerNorm input
 * gradient, dbias (nullable) its column sums; upstream = dout (fp32) + dout_bf16, either may be NULL.
 * workspace: xr_enc_ln_workspace_bytes(1).                                                                  */
int xr_enc_add_ln_fwd(const void* y, int y_dtype, const float* bias, const float* residual, const float* gamma,
                      const float* beta, int64_t n_tok, int64_t dim, float eps, float* out, void* out_bf16,
                      float* stats, const int64_t* rng, float drop_p, int site, void* stream);
int xr_enc_add_ln_bwd(const void* y, int y_dtype, const float* bias, const float* residual, const float* gamma,
                      const float* stats, const float* dout, const void* dout_bf16, int64_t n_tok, int64_t dim,
                      float* dresidual, void* dy, float* dbias, float* dgamma, float* dbeta, void* workspace,
                      const int64_t* rng, float drop_p, int site, void* stream);
/* out[c] = sum over rows of x[r][c] (a linear layer's bias gradient: torch.nn.Linear inside BertSelfAttention /
 * BertIntermediate); fp32 accumulation in a fixed order.  width must be a multiple of 16 bytes of elements. */
size_t xr_enc_colsum_workspace_bytes(int64_t width);
int xr_enc_colsum(const void* x, int dtype, int64_t rows, int64_t width, float* out, void* workspace, void* stream);
/* exact (erf) GELU: dy == NULL -> out = gelu(x); else out = dy * gelu'(x).                               */
int xr_enc_gelu(const void* x, const void* dy, int64_t n, int dtype, void* out, void* stream);
/* causal + key-padding multi-head self-attention on qkv (B, L, 3 * n_heads * 32) = [Q | K | V]; keymask
 * (B, L) bytes.  dctx == NULL: forward, out = ctx (B, L, n_heads * 32), lse (B, n_heads, L) written.
 * dctx != NULL: backward, out = dqkv, lse and ctx from the forward.  Deterministic (no atomics).          */
int xr_enc_attention(const void* qkv, const uint8_t* keymask, const void* ctx, const void* dctx, float* lse,
                     int64_t batch, int64_t seq_len, int64_t n_heads, int64_t head_dim, int dtype, void* out,
                     const int64_t* rng, float drop_p, int site, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XFMR_B200_H */
