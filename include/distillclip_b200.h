/*
 * distillclip_b200 -- C ABI of the B200-native distillation-loss hot path.
 *
 * The reference (ForJadeForest/DistillCLIP) is pure Python: its "operator interface" for this
 * path is the nn.Module API of model/_loss.py and model/loss_component/ (SURVEY.md section 8b).
 * Every entry point below is what a ctypes/cffi binding inside those modules' forward()/backward()
 * would call; each cites the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain pointers and sizes only; all buffers are caller-allocated DEVICE memory unless the
 *     parameter says "host"; pointer *arrays* (`const void* const*`) are HOST arrays of device pointers;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); every call only
 *     enqueues work on it, never synchronises and never allocates, so calls are CUDA-graph capturable;
 *   - return 0 on success, non-zero on error; `dcb_last_error()` returns a thread-local message;
 *   - dtypes: DCB_BF16 / DCB_F16 / DCB_F32 for student+teacher inputs (both the same), and for
 *     gradient outputs either the input dtype or DCB_F32;
 *   - loss values are produced in two steps so that results are deterministic (no float atomics):
 *     a kernel writes per-CTA partial sums (double) and `dcb_finalize` reduces them in a fixed order.
 */
#ifndef DISTILLCLIP_B200_H_
#define DISTILLCLIP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { DCB_BF16 = 0, DCB_F16 = 1, DCB_F32 = 2 };
enum { DCB_MAX_LAYERS = 16, DCB_MAX_PARTIALS = 4096, DCB_MAX_TERMS = 16 };

/* library / build identification */
int dcb_version(void);
const char* dcb_last_error(void);
/* compute capability the kernels were compiled for (100 = sm_100a) */
int dcb_compiled_arch(void);

/* ---------------------------------------------------------------------------------------------
 * Hidden-state / embedding MSE, forward and backward in ONE pass.
 * Replaces HiddenMSE.forward (model/loss_component/hidden_mse.py:9-17), EmbedMSELoss.forward
 * (model/loss_component/embed_mse.py:9-10) and the autograd backward of nn.MSELoss under them.
 *   value  = (1/divisor) * sum_l mean((stu_l - tea_l)^2)          (divisor = len(stu_hidden), :16)
 *   grad_l = grad_scale * 2 (stu_l - tea_l) / (numel_l * divisor)  (written when grad_stu[l] != NULL)
 * partials: >= DCB_MAX_PARTIALS doubles; *n_partials (host) receives how many were written.
 * --------------------------------------------------------------------------------------------- */
int dcb_mse_fwd_bwd(int n_layers, const void* const* stu, const void* const* tea, void* const* grad_stu,
                    const int64_t* numel, int in_dtype, int grad_dtype, int divisor, float grad_scale,
                    double* partials, int* n_partials, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Attention-map KL, forward and backward in ONE pass.
 * Replaces AttentionProbsKL.forward (model/loss_component/attention_probs_kl.py:10-22).
 * Layer l: stu_l is [batch, stu_heads, positions], tea_l is [batch, tea_heads, positions]
 * (positions = N*N, contiguous).  s = head-mean(stu), t = head-mean(tea):
 *   value    = (1/divisor) * sum_l sum_{b,p} [xlogy(t,t) - t log s]          (KLDivLoss 'sum', :8)
 *   grad_l   = -grad_scale * t / (s * stu_heads * divisor), the same for every head
 * NaN/Inf propagate exactly as in the reference (both maps zero at one position -> NaN).
 * --------------------------------------------------------------------------------------------- */
int dcb_attn_kl_fwd_bwd(int n_layers, const void* const* stu, const void* const* tea, void* const* grad_stu,
                        const int64_t* batch, const int32_t* stu_heads, const int32_t* tea_heads,
                        const int64_t* positions, int in_dtype, int grad_dtype, int divisor, float grad_scale,
                        double* partials, int* n_partials, void* stream);

/* ---------------------------------------------------------------------------------------------
 * ALL streaming losses of one tower in ONE launch, including the weighting below.
 * Replaces LossCalculator.cal_one_tower_loss's loops (model/_loss.py:155-202) over embedding_mse, hidden_rep_mse and
 * attention_probs_kl and their autograd backward.  Segment k (one layer of one loss term) has
 *   kind[k]  0 = MSE (numel[k]),  1 = attention KL (batch[k], stu_heads[k], tea_heads[k], positions[k]),
 *            2 = L1 (out_l1.py, numel[k]),  3 = cosine rows (out_cos.py; batch[k] rows of positions[k] elements),
 *            4 = MSE of head-mean maps (attention_probs_mse.py / attention_score_mse.py; shapes as kind 1);
 *   term[k]  index of the loss term it adds to (< n_terms <= 8, segments grouped by term); divisor[k], grad_scale[k]
 *            as in the calls above.
 * out[q] = scale[q] * value_q (q < n_terms), out[n_terms] = sum_q percent[q] * out[q]  (_loss.py:199-200).
 * partials: n_terms rows of partial_stride (>= dcb_tower_grid()) doubles of scratch; ticket: one uint32, zero before the
 * first launch (the kernel resets it).  Values are reduced in a fixed order by the last CTA to finish: deterministic,
 * no extra launch.  Terms in ext_mask were already reduced to ext_count partials per row by another kernel on the same
 * stream (n_seg may then be 0: this call only finishes the weighting); pass 0 / 0 otherwise.
 * --------------------------------------------------------------------------------------------- */
int dcb_tower_grid(void);
int dcb_tower_fwd_bwd(int n_seg, const int32_t* kind, const int32_t* term, const void* const* stu,
                      const void* const* tea, void* const* grad_stu, const int64_t* numel,
                      const int64_t* batch, const int32_t* stu_heads, const int32_t* tea_heads,
                      const int64_t* positions, const int32_t* divisor, const float* grad_scale, int n_terms,
                      const float* scale, const float* percent, int in_dtype, int grad_dtype,
                      double* partials, int partial_stride, uint32_t ext_mask, int ext_count,
                      uint32_t* ticket, float* out, const float* fwd_mult, const float* const* upstream,
                      const float* up_mult, const float* expected, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Deterministic reduction + weighting (model/_loss.py:195-200 and :148-152).
 *   out[k]        = scale[k] * sum(partials[k][0 .. counts[k]))            k < n_terms
 *   out[n_terms]  = sum_k percent[k] * out[k]
 * partials: host array of device pointers; counts/scale/percent: host arrays; out: device floats.
 * --------------------------------------------------------------------------------------------- */
int dcb_finalize(int n_terms, const double* const* partials, const int32_t* counts, const float* scale,
                 const float* percent, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused global-batch contrastive (InfoNCE) + teacher/student logit KL from EMBEDDINGS.
 * Replaces, without ever writing the B x B logits:
 *   CLIPModel.forward's normalise + `image_feature @ text_feature.t()` (model/component/clip_model.py:36-44),
 *   HardLabel.forward (model/loss_component/hard_label.py:10-12) on i2t and t2i,
 *   SoftLabel.forward (model/loss_component/soft_label.py:11-16) on i2t and t2i,
 *   the 0.5*(i2t + t2i) combination of model/_loss.py:130-137, and autograd through all of them.
 *
 * Row-sharded: this process owns `rows_local` rows [row_offset, row_offset+rows_local) of the global
 * batch on the "a" side and sees all `cols` rows of the "b" side (after an all-gather, or the local
 * batch when not sharded).  A "direction" is (a=image,b=text) for i2t rows or (a=text,b=image) for t2i.
 * Embeddings are bf16 or fp16 (`dtype`), row-major [rows, dim] with dim % 8 == 0, 16-byte aligned.
 * Cosine logits are bounded by 1, so softmax uses the fixed shift 1 and partial sums over column
 * ranges / ranks simply add.  temperature must be >= 0.025 on this path (exp((S-1)/T) stays normal).
 * --------------------------------------------------------------------------------------------- */

/* inv_norm[k][i] = 1 / ||mats[k][i,:]||_2  (clip_model.py:37-38).  dtype: DCB_BF16 / DCB_F16 / DCB_F32. */
int dcb_row_inv_norm(int n_mats, const void* const* mats, float* const* inv_norm, const int64_t* rows,
                     int64_t dim, int dtype, void* stream);

/* out[d][j] = fp16(in[j][d] * inv_norm[j]); out is [dim, out_pitch_elems] (pitch >= rows, % 8 == 0, pad columns zeroed).
 * The gradient GEMM wants the normalised student b-side K-major, so it is transposed once per backward. */
int dcb_transpose_norm_f16(const void* in, const float* inv_norm, void* out, int64_t rows, int64_t dim,
                           int64_t out_pitch_elems, int dtype, void* stream);

/* dcb_clip_row_stats: ONE fused tcgen05 kernel for one direction.  For each local row i:
 *   stats[0][i] = sum_j exp(S_ij - 1)                 (hard label; S = student cosine logits)
 *   stats[1][i] = Q_i = sum_j [es_ij - et_ij + et_ij (T_ij - S_ij)/T],  es = exp((S-1)/T), et = exp((T-1)/T): the
 *                 second-order part of Zs_i - Zt_i (Zs = Zt + Q - W/T), carried instead of Zs so that the KL of a row,
 *                 -m + log1p(m + Q/Zt) with m = -W/(T Zt), keeps fp32 rounding relative to its own (small) size
 *   stats[2][i] = Zt_i = sum_j exp((T_ij - 1)/T)
 *   stats[3][i] = W_i = sum_j exp((T_ij - 1)/T) (T_ij - S_ij)      stats[4][i] = S_ii (global diagonal)
 * tea_* may all be NULL (hard label only; stats[1..3] are then 0).  stats: [5, rows_local] floats.
 * rowloss: [2, rows_local] doubles = {CE_i = 1 + log stats0 - stats4,  KL_i / T^2 = stats3/(T stats2) + log(stats1/stats2)}.
 * col_stats (optional, [4, cols] floats): the SAME four sums taken down the columns over this call's rows, i.e. this
 *   call's contribution to the row statistics of the opposite direction (t2i when a = image) -- one pass over the logits
 *   serves both directions; all-reduce over ranks when sharded, then dcb_clip_col_finish.  NULL = row statistics only.
 * workspace: dcb_clip_workspace_bytes(rows_local, cols) bytes of device scratch.
 * dump_s / dump_t: optional [rows_local, cols] fp32 buffers that receive the logits (tests only; NULL in production). */
int64_t dcb_clip_workspace_bytes(int64_t rows_local, int64_t cols);
int dcb_clip_row_stats(const void* stu_a, const void* stu_b, const void* tea_a, const void* tea_b,
                       const float* stu_a_inv, const float* stu_b_inv, const float* tea_a_inv, const float* tea_b_inv,
                       int64_t rows_local, int64_t row_offset, int64_t cols, int64_t dim, int dtype, float temperature,
                       float* stats, double* rowloss, float* col_stats, void* workspace, float* dump_s, float* dump_t,
                       void* stream);

/* Retrieval rank of the label (validation metrics, model/dual_distill_model.py:129-187,220-224: norm_and_logits + top-k
 * accuracy): same tiles as dcb_clip_row_stats without a teacher, and stats[1][i] = #{j : S_ij > ref[i]} (exact integer in a
 * float).  With ref = stats[4] (= S_ii) of a previous dcb_clip_row_stats call the label is in the top k iff stats[1][i] < k;
 * the diagonal itself is produced by the same instructions and never counts as greater than itself. */
int dcb_clip_rank_counts(const void* a, const void* b, const float* a_inv, const float* b_inv, int64_t rows_local,
                         int64_t row_offset, int64_t cols, int64_t dim, int dtype, const float* ref, float* stats,
                         double* rowloss, void* workspace, void* stream);

/* Row statistics / per-row losses of the opposite direction for this rank's rows from the complete column statistics:
 * stats[k][i] = col_stats[k][row_offset + i] (k < 4), stats[4][i] = diag_local[i] (= stats[4] of dcb_clip_row_stats:
 * the logit matrix has one diagonal), rowloss as above. */
int dcb_clip_col_finish(const float* col_stats, int64_t cols_total, const float* diag_local, int64_t row_offset,
                        int64_t rows_local, float temperature, int has_teacher, float* stats, double* rowloss, void* stream);

/* Loss values of both directions from this rank's per-row losses (the `rowloss` outputs of dcb_clip_row_stats):
 *   sums[0..3] (double) = {sum_i CE_i (i2t), sum_i CE_i (t2i), T^2 sum_i KL_i (i2t), T^2 sum_i KL_i (t2i)}
 *   out[0] = 0.5 (sums[0] + sums[1]) / global_batch   (hard_label.py:12 'mean', _loss.py:131)
 *   out[1] = 0.5 (sums[2] + sums[3])                  (soft_label.py:8 'sum', _loss.py:135-136)
 * When sharded, all-reduce `sums` over ranks and rescale instead of using `out`. */
int dcb_clip_losses(const double* rowloss_i2t, const double* rowloss_t2i, int64_t rows_i2t, int64_t rows_t2i,
                    int64_t global_batch, float temperature, int has_teacher, double* sums, float* out, void* stream);

/* coef[0][i] = gh/(2 B A_i), coef[1][i] = gs T/(2 Zs_i) with Zs = Zt + Q - W/T, coef[2][i] = gs T/(2 Zt_i) from stats [5, rows];
 * upstream: device float[2] = {gh = d total / d hard, gs = d total / d soft}.
 * gmax: device float[1], receives max_i (|coef0| + |coef1| + |coef2|) -- the bound that fixes the fp16 scale of G. */
int dcb_clip_grad_coef(const float* stats, int64_t rows, int64_t global_batch, float temperature, int has_teacher,
                       const float* upstream, float* coef, float* gmax, void* stream);

/* dcb_clip_row_grads: backward for one direction, recomputing the logits tile by tile:
 *   G_ij = e1_ij (coef_row[0][i] + coef_col[0][j]) + es_ij (coef_row[1][i] + coef_col[1][j])
 *        - et_ij (coef_row[2][i] + coef_col[2][j])
 *   acc_parts[s][i,:] = partial sums over column range s of  2^k sum_j G_ij b_hat_j      (fp32)
 * coef_col are the coefficients of the OPPOSITE direction for all `cols` rows (the column softmax of the same logits);
 * gmax_row / gmax_col the matching outputs of dcb_clip_grad_coef (2^k = largest power of two with
 * 2^k (gmax_row + gmax_col) <= 2^14: G goes through the tensor cores as fp16).
 * stu_b_t: dcb_transpose_norm_f16(stu_b).  acc_parts: dcb_clip_grad_workspace_bytes(...) bytes,
 * dcb_clip_grad_splits(...) partial buffers of [rows_local, dim]. */
int64_t dcb_clip_grad_workspace_bytes(int64_t rows_local, int64_t cols, int64_t dim);
int dcb_clip_grad_splits(int64_t rows_local, int64_t cols, int64_t dim);
int dcb_clip_row_grads(const void* stu_a, const void* stu_b, const void* tea_a, const void* tea_b,
                       const void* stu_b_t, int64_t bt_pitch_elems,
                       const float* stu_a_inv, const float* stu_b_inv, const float* tea_a_inv, const float* tea_b_inv,
                       const float* coef_row, const float* coef_col, const float* gmax_row, const float* gmax_col,
                       int64_t rows_local, int64_t cols, int64_t dim, int dtype, float temperature,
                       float* acc_parts, void* stream);

/* Same contract as dcb_clip_row_grads, CTA-pair implementation (tcgen05 cta_group::2, clusters of 2): the whole
 * embedding dimension is accumulated in one pass, so the logits are recomputed once per direction.  dim <= 768.
 * acc_parts: dcb_clip_pair_splits(...) buffers of [rows_local, dim].  dump_s: tests only (NULL in production).
 * trace: profiling only, 2 x 64 x 16 int64 clock64() stamps of the first cluster's pipeline events (NULL in production). */
int dcb_clip_pair_supported(int64_t dim);
int dcb_clip_pair_splits(int64_t rows_local, int64_t cols, int64_t dim);
int dcb_clip_row_grads_pair(const void* stu_a, const void* stu_b, const void* tea_a, const void* tea_b,
                            const void* stu_b_t, int64_t bt_pitch_elems,
                            const float* stu_a_inv, const float* stu_b_inv, const float* tea_a_inv, const float* tea_b_inv,
                            const float* coef_row, const float* coef_col, const float* gmax_row, const float* gmax_col,
                            int64_t rows_local, int64_t cols, int64_t dim, int dtype, float temperature,
                            float* acc_parts, void* g_out, int64_t g_pitch_elems, float* dump_s, long long* trace,
                            void* stream);

/* Single-recompute backward: when g_out != NULL the pair kernel also stores its scaled fp16 gradient tiles
 * G[i, j] 2^k = dL/dS_ij 2^k ([rows_local, g_pitch_elems], pitch >= cols, multiple of 8) and the b-side accumulator
 *     acc_b[j, :] = sum_i G[i, j] a_hat[i, :]
 * comes from one tcgen05 GEMM over the stored tiles (A = G^T read as an MN-major operand) instead of a second
 * recompute of the logits.  a_hat_t: [dim, at_pitch_elems] fp16 from dcb_transpose_norm_f16 of the a side.
 * acc_parts: dcb_clip_gt_splits(...) buffers of [cols, dim] fp32 for dcb_clip_grad_finish (same 2^k scale). */
int dcb_clip_gt_splits(int64_t rows, int64_t cols, int64_t dim);
int dcb_clip_gt_splits_scatter(int64_t rows, int64_t cols, int64_t dim);   /* K splits dcb_clip_col_grads_scatter uses */
int dcb_clip_col_grads_from_g(const void* g, int64_t g_pitch_elems, const void* a_hat_t, int64_t at_pitch_elems,
                              int64_t rows, int64_t cols, int64_t dim, float* acc_parts, void* stream);

/* The same GEMM fused with the gradient reduce-scatter of the row-sharded path: output row j belongs to rank
 * j / (cols / n_dest) and is stored from the epilogue straight into that rank's buffer dest_parts[rank] (peer-mapped device
 * memory, NVLink stores), slot src_slot * k_split + ks of [n_dest * dcb_clip_gt_splits(...)][cols / n_dest][dim] fp32.  After
 * a cross-rank barrier the owner reduces its slots in a fixed order with dcb_clip_grad_finish (n_split = n_dest * k_split):
 * no NCCL reduce-scatter, no atomics, deterministic. */
int dcb_clip_col_grads_scatter(const void* g, int64_t g_pitch_elems, const void* a_hat_t, int64_t at_pitch_elems,
                               int64_t rows, int64_t cols, int64_t dim, void* const* dest_parts, int n_dest, int src_slot,
                               void* stream);

/* grad_a[i,:] = r_i (acc_i - a_hat_i (a_hat_i . acc_i)),  acc_i = 2^-k sum_s acc_parts[s][i,:] - (gh/B) b_hat_{row_offset+i}
 * (the -[i==j] label term of the cross entropy, added here in fp32, then the x/||x|| Jacobian of clip_model.py:37-38). */
int dcb_clip_grad_finish(const float* acc_parts, int n_split, const void* stu_a, const float* stu_a_inv,
                         const void* stu_b, const float* stu_b_inv, int64_t rows, int64_t cols, int64_t dim,
                         int64_t row_offset, int64_t global_batch, const float* upstream, const float* gmax_row,
                         const float* gmax_col, int in_dtype, void* grad_a, int grad_dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused contrastive PIPELINE (distillclip_b200/pipeline.py): the same tcgen05 kernels driven by four small fused kernels,
 * 7 launches per fwd+bwd on one GPU, and on N GPUs three peer-memory exchanges with one barrier each instead of collectives.
 * Replaces, for the GLOBAL batch: clip_model.py:36-44 (normalise + matmul), hard_label.py:10-12, soft_label.py:11-16, the
 * 0.5 (i2t + t2i) sums and the scale / percent weighting of _loss.py:130-137,231-234, and autograd through all of it.
 * --------------------------------------------------------------------------------------------- */
/* stream-ordered device-to-device copy (peer-mapped source: a copy-engine pull over NVLink) */
int dcb_memcpy_async(void* dst, const void* src, int64_t bytes, void* stream);
/* up to 64 (dst, src, bytes) copies in ONE kernel launch (16-byte aligned; peer-mapped sources = loads over NVLink): the
 * latency-optimised exchange for small global batches */
int dcb_peer_gather(int n_copies, void* const* dst, const void* const* src, const int64_t* bytes, void* stream);

/* prep: for up to 4 matrices [rows, dim] (bf16/fp16, 16-byte aligned): inv_norm[k][i] = 1/||x_i|| (clip_model.py:37-38);
 * copy_out[k] (optional) = the raw rows (this rank's slice of the buffer peers pull from); tr_out[k] (optional) = fp16
 * [dim, tr_pitch_elems[k]] = (x * inv_norm).T, the K-major operand of the gradient GEMMs. */
int dcb_clip_prep(int n_mats, const void* const* mats, float* const* inv_norm, void* const* copy_out, void* const* tr_out,
                  const int64_t* tr_pitch_elems, int64_t rows, int64_t dim, int dtype, void* stream);

/* Similarity tiles of the local rows against ONE chunk of the columns (the text rows of one source rank): partial row sums
 * into ws_chunk[dcb_clip_fwd_chunk_parts(...)][4][rows], S_ii into diag[rows] where the label column
 * (label_col0 + i, relative to the chunk) falls inside the chunk, column sums into
 * col_part_chunk[ceil(rows/128)][4][col_part_ld] (pointer advanced to the chunk's first column). */
int dcb_clip_fwd_chunk_parts(int64_t rows_local, int64_t cols_chunk);
int dcb_clip_fwd_chunk(const void* stu_a, const void* stu_b, const void* tea_a, const void* tea_b,
                       const float* stu_a_inv, const float* stu_b_inv, const float* tea_a_inv, const float* tea_b_inv,
                       int64_t rows_local, int64_t label_col0, int64_t cols_chunk, int64_t dim, int dtype, float temperature,
                       float* ws_chunk, float* diag, float* col_part_chunk, int64_t col_part_ld, float* ws_extra_chunk,
                       float* diag_t, void* stream);
/* ws_extra_chunk (optional, [parts][2][rows]) and diag_t[rows]: ALSO the row sums of relu(S_ij - T_ij) and (S_ij - T_ij)^2 and
 * the teacher diagonal T_ii -- CLIPCosDiff (clip_cos_diff.py:5-23) and LogitsMSE (logits_mse.py:9-10) from the same tiles, so the
 * shipped stage-3 recipe (config/final_config/l_clip.yaml:30) never materialises the B x B logits either. */

/* Statistics slot exchanged between ranks, in floats: [4][cols] column sums | [rows_per_rank] S_ii | 2 doubles
 * {sum CE_i2t, sum KL_i2t / T^2, sum relu(T_ii - S_ii), sum_{i != j} relu(S_ij - T_ij), sum (S_ij - T_ij)^2} | 4 floats {max of
 * the three unit coefficients, 0} | 2 floats padding. */
int64_t dcb_clip_slot_floats(int64_t cols, int64_t rows_per_rank);
int64_t dcb_clip_post_scratch_bytes(int64_t rows, int64_t cols);   /* zero-initialised once; the kernels reset their ticket */

/* post1: row statistics stats[5][rows] of the local rows (fixed-order sum of the n_part partial sets), UNIT gradient
 * coefficients coef_row[4][rows] = {1/(2 B A_i), T/(2 Zs_i), T/(2 Zt_i), [T_ii > S_ii]} (no upstream gradient in them; the
 * fourth row is the diagonal flag of CLIPCosDiff, 0 without ws_extra / diag_t), and this rank's slot
 * (column sums over its rows, S_ii, i2t loss sums, maxima) stored into dest_slots[0..n_dest): this rank's slot inside every
 * rank's slot buffer (peer-mapped addresses: NVLink stores; n_dest = 1 with a local buffer when a collective follows). */
int dcb_clip_post1(const float* ws, int n_part, const float* diag, const float* col_part, int row_blocks, int64_t rows,
                   int64_t cols, float temperature, int has_teacher, int64_t global_batch, float* stats, float* coef_row,
                   void* const* dest_slots, int n_dest, const float* ws_extra, const float* diag_t, void* scratch, void* stream);

/* post2 (after the cross-rank barrier): slots[n_src][dcb_clip_slot_floats] summed in source order -> col_stats[4][cols],
 * coef_col[3][cols] (unit coefficients of the t2i direction), bounds[6] = maxima over all rows / all columns (they fix the
 * fp16 scale of the gradient tiles), out[9] = {hard, soft, hard s_hard, soft s_soft, total, cos_diff, logits_mse, cos_diff s_cos,
 * logits_mse s_mse}, total = p_hard out[2] + p_soft out[3] + p_cos out[7] + p_mse out[8], with hard = 0.5 (CE_i2t + CE_t2i) (mean),
 * soft = 0.5 T^2 (KL_i2t + KL_t2i) (sum), cos_diff = mean relu(T_ii - S_ii) + mean_{i != j} relu(S_ij - T_ij), logits_mse =
 * mean (S - T)^2 of the GLOBAL batch (_loss.py:130-145,231-234).  weights8 (host) = {p_hard, p_soft, s_hard, s_soft, p_cos,
 * p_mse, s_cos, s_mse}. */
int dcb_clip_post2(const float* slots, int n_src, int64_t rows_per_src, int64_t cols, float temperature, int has_teacher,
                   int64_t global_batch, const float* weights8, float* col_stats,
                   float* coef_col, float* bounds, float* out, void* scratch, void* stream);

/* Backward, first kernel: as dcb_clip_row_grads_pair, but with UNIT coefficients multiplied on the device by
 *   up_hard = *g_total * w_hard + *g_hard * s_hard,  up_soft = *g_total * w_soft + *g_soft * s_soft
 * (g5 = {g_total, g_hard, g_soft, g_cos, g_mse}: 0-dim fp32 device scalars from autograd, NULL = no gradient for that output;
 * w8 (host) = {w_hard, w_soft, s_hard, s_soft, w_cos, w_mse, s_cos, s_mse}, w = percent * scale), the fp16 tile scale from
 * bounds[6], and stu_b_t stored as one [dim, bt_block_cols] block per source rank ([cols / bt_block_cols][dim][pitch]).
 * extra = 1 adds d(CLIPCosDiff)/dS_ij = up_cos [S_ij > T_ij] / (B (B - 1)) (i != j; row_offset = global index of local row 0)
 * and d(LogitsMSE)/dS_ij = 2 up_mse (S_ij - T_ij) / B^2 to the tiles. */
int dcb_clip_pair_bwd(const void* stu_a, const void* stu_b, const void* tea_a, const void* tea_b,
                      const void* stu_b_t, int64_t bt_pitch_elems, int64_t bt_block_cols,
                      const float* stu_a_inv, const float* stu_b_inv, const float* tea_a_inv, const float* tea_b_inv,
                      const float* coef_row, const float* coef_col, const float* bounds,
                      const float* const* g5, const float* w8, int extra, int64_t row_offset, int64_t global_batch,
                      int64_t rows_local, int64_t cols, int64_t dim, int dtype,
                      float temperature, float* acc_parts, void* g_out, int64_t g_pitch_elems, void* stream);

/* SPLIT backward (default for large batches): the logits are recomputed once by a kernel of the forward's shape (256 rows
 * per CTA pair, double-buffered S/T accumulators) that only STORES the scaled fp16 gradient tiles G 2^k (arguments as
 * dcb_clip_pair_bwd, without accumulators); both towers' gradients are then tcgen05 GEMMs over the stored tiles:
 * dcb_clip_row_grads_from_g (acc_a = G b_hat, A = G K-major; b_hat_t in blocks of bt_block_cols columns per source rank) and
 * dcb_clip_col_grads_from_g / _scatter (acc_b = G^T a_hat).  acc_parts: dcb_clip_rg_splits(...) buffers of [rows_local, dim].
 * Replaces autograd through clip_model.py:36-44 + hard_label.py:10-12 + soft_label.py:11-16 (+ clip_cos_diff.py:16-23,
 * logits_mse.py:9-10 with extra = 1), like the pair kernel it is an alternative to. */
int dcb_clip_g_tiles(const void* stu_a, const void* stu_b, const void* tea_a, const void* tea_b,
                     const float* stu_a_inv, const float* stu_b_inv, const float* tea_a_inv, const float* tea_b_inv,
                     const float* coef_row, const float* coef_col, const float* bounds,
                     const float* const* g5, const float* w8, int extra, int64_t row_offset, int64_t global_batch,
                     int64_t rows_local, int64_t cols, int64_t dim, int dtype, float temperature, void* g_out,
                     int64_t g_pitch_elems, void* stream);
int dcb_clip_rg_splits(int64_t rows_local, int64_t cols, int64_t dim);
int dcb_clip_row_grads_from_g(const void* g, int64_t g_pitch_elems, const void* b_hat_t, int64_t bt_pitch_elems,
                              int64_t bt_block_cols, int64_t rows_local, int64_t cols, int64_t dim, float* acc_parts,
                              void* stream);

/* Backward, last kernel, both towers in one launch (side a = image rows from the pair kernel's accumulators, side b = text
 * rows from the G^T GEMM's): grad[i,:] = r_i (acc_i - x_hat_i (x_hat_i . acc_i)), acc_i = 2^-k sum_s acc[s][i,:] - (up_hard / B)
 * y_hat_{label_offset + i}.  A side with grad == NULL is skipped.  cos_flag (optional, [rows] = coef_row row 3): the diagonal
 * term of CLIPCosDiff, -up_cos [T_ii > S_ii] / B, joins the label term (and the tiles carried the extra terms). */
int dcb_clip_finish2(const float* acc_a, int n_split_a, int64_t split_stride_a, const void* a, const float* a_inv, void* grad_a,
                     int64_t rows_a, const void* a_label, const float* a_label_inv, int64_t a_label_rows, int64_t a_label_offset,
                     const float* acc_b, int n_split_b, int64_t split_stride_b, const void* b, const float* b_inv, void* grad_b,
                     int64_t rows_b, const void* b_label, const float* b_label_inv, int64_t b_label_rows, int64_t b_label_offset,
                     int64_t dim, int64_t global_batch, const float* const* g5, const float* w8, const float* cos_flag,
                     const float* bounds, int in_dtype, int grad_dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Per-module API on MATERIALISED logits (HardLabel / SoftLabel keep their logits signature).
 * logits: [n, n] with element strides (row_stride, col_stride) so that `logits.T` views work.
 * mode 0 = HardLabel (hard_label.py:10-12), mode 1 = SoftLabel (soft_label.py:11-16),
 * mode 2 = CLIPCosDiff (clip_cos_diff.py:5-23; rowloss already carries the 1/n and 1/(n(n-1)) means, reduce with scale 1).
 *   saved[i]   = {max_s, Z_s, max_t, Z_t} (float4 per row, for the backward)
 *   rowloss[i] = CE_i (mode 0) or KL_i without the T^2 factor (mode 1), double; reduce with dcb_finalize
 *                (scale 1/n resp. T^2).
 *   grad_logits: contiguous [n, n], dtype `grad_dtype`; upstream: device float[1].
 * --------------------------------------------------------------------------------------------- */
int dcb_logits_row_stats(const void* stu_logits, int64_t stu_rs, int64_t stu_cs,
                         const void* tea_logits, int64_t tea_rs, int64_t tea_cs,
                         int64_t n, int dtype, float temperature, int mode, float* saved, double* rowloss, void* stream);
int dcb_logits_row_grads(const void* stu_logits, int64_t stu_rs, int64_t stu_cs,
                         const void* tea_logits, int64_t tea_rs, int64_t tea_cs,
                         int64_t n, int dtype, float temperature, int mode, const float* saved,
                         const float* upstream, void* grad_logits, int grad_dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Row-softmax losses between pooled outputs [rows, cols] (contiguous):
 *   mode 0 = OutKLLoss (model/loss_component/out_kl.py:12-16): KLDiv(sum)(log_softmax(s/T), softmax(t/T)) * T^2
 *   mode 1 = OutCELoss (model/loss_component/out_ce.py:9-13):  CrossEntropy(mean)(s, softmax(t))
 * saved: float4 per row {max_s, Z_s, max_t, Z_t}; rowloss: double per row (reduce with dcb_finalize, scale T^2 resp.
 * 1/rows); grads: g = up T (p^s - p^t) resp. up (softmax(s) - p^t) / rows; upstream: device float[1].
 * --------------------------------------------------------------------------------------------- */
int dcb_row_softmax_stats(const void* stu, const void* tea, int64_t rows, int64_t cols, int dtype, float temperature, int mode,
                          float* saved, double* rowloss, void* stream);
int dcb_row_softmax_grads(const void* stu, const void* tea, int64_t rows, int64_t cols, int dtype, float temperature, int mode,
                          const float* saved, const float* upstream, void* grad, int grad_dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LastValueMapKL (model/loss_component/last_value_map_kl.py:10-14): KLDiv(sum)(softmax(stu, dim=1).log(), softmax(tea, dim=1))
 * on [batch, heads, positions] value-relation maps; the softmax runs over the HEAD axis.  One pass, forward value and
 * student gradient (p^s - p^t) * grad_scale * (*fwd_mult); heads <= 16.  partials: >= DCB_MAX_PARTIALS doubles, reduce with
 * dcb_finalize.  fwd_mult (optional device scalar): the AMP GradScaler's scale.  Regrad mode (upstream != NULL, partials
 * NULL): gradients only, coefficient grad_scale * (*upstream), skipped entirely when *upstream == expected * (*fwd_mult).
 * --------------------------------------------------------------------------------------------- */
int dcb_value_map_kl_fwd_bwd(const void* stu, const void* tea, void* grad_stu, int64_t batch, int64_t heads,
                             int64_t positions, int in_dtype, int grad_dtype, float grad_scale, double* partials,
                             int* n_partials, const float* fwd_mult, const float* upstream, float expected, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DISTILLCLIP_B200_H_ */
