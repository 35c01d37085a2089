/* kd_b200.h - C ABI of libkd_b200.so, the B200 (sm_100a) implementation of the
 * speech-distill knowledge-distillation hot path.
 *
 * The reference (indiejoseph/speech-distill) is pure Python and has no FFI of its own; the
 * entry points below are what a binding for its one arithmetic module would call.  Each one
 * cites the reference code it replaces (paths relative to the reference repo).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the comment says "host";
 *   - tensors are row-major, last dimension contiguous; strides are in ELEMENTS;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - every call is asynchronous on `stream`, never synchronises the host and never allocates;
 *   - return value 0 = success, non-zero = error, text via kd_last_error() (thread local);
 *   - dtype codes: KD_DTYPE_F32 / KD_DTYPE_BF16 / KD_DTYPE_F16.
 *
 * Row convention (distillation_loss.py:31-45): with logits [B,T,V] and labels [B,T], row (b,t)
 * is scored against labels[b,t+1]; it is VALID iff t < T-1, labels[b,t+1] != ignore_index and
 * (mask == NULL or mask[b,t+1] != 0).  N = number of valid rows.
 *
 * Sums record (float[8], written by the *_fwd* calls):
 *   [0] sum over valid rows of CE_r  = LSE(z) - z[label]                    (distillation_loss.py:123)
 *   [1] sum over valid rows of KL_r  (temperature tau, WITHOUT the tau^2)   (:66-68 dense, :94-106 sparse)
 *   [2] dense : sum of teacher CE_r = LSE(y) - y[label]                     (:71)
 *       sparse: sum of v[r,k] over hits (i[r,k] == label)                   (:110-116)
 *   [3] N (valid rows seen by this call)
 *   [4] sparse: number of hits; dense: 0
 *   [5..7] reserved (0)
 * Normalisation (/N, *tau^2, alpha mix) is a separate call so that a data-parallel caller can
 * all-reduce the record first.
 */
#ifndef KD_B200_H_
#define KD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KD_DTYPE_F32 0
#define KD_DTYPE_BF16 1
#define KD_DTYPE_F16 2

#define KD_ABI_VERSION 5

/* OR-ed into grad_dtype of the fused backward: dH is written as fp32 (a vocab-parallel caller sums the
 * per-slice partial dH across ranks before rounding) while dW keeps the base dtype. */
#define KD_GRAD_DH_F32 0x100

/* Teacher kinds for the fused LM-head entry points. */
#define KD_TEACHER_NONE 0   /* CE only (stage1 warm-up, stage1.py:298-340) */
#define KD_TEACHER_DENSE 1  /* full-vocab teacher logits (distillation_loss.py:56-71) */
#define KD_TEACHER_SPARSE 2 /* top-k log-probs + indices (distillation_loss.py:73-118) */

int kd_version(void);
const char* kd_last_error(void); /* host string, valid until the next failing call on this thread */

/* Number of SMs / compute capability of the current device (host helper for callers that size work). */
int kd_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- row bookkeeping -------------------------------------------------------------------
 * distillation_loss.py:34-45: builds, for every row r = b*T + t, row_target[r] = labels[b,t+1]
 * if the row is valid else -1, and counts the valid rows.  labels int64 [B,T] contiguous;
 * mask uint8 [B,T] (non-zero = keep) or NULL.  row_target may be NULL (count only). */
int kd_prepare_rows(const int64_t* labels, const uint8_t* mask, int B, int T, int64_t ignore_index,
                    int32_t* row_target, int32_t* n_valid, void* stream);

/* distillation_loss.py:68 (*tau^2, batchmean /N), :116-118, :123, :126.
 * sums float[8] as above -> losses float[4] = (total, task, distill, teacher_task).
 * N == 0 gives four zeros (:47-53). */
int kd_finalize_losses(const float* sums, float tau, float alpha, int sparse, float* losses, void* stream);

/* ---- K2: streaming KD on materialised logits --------------------------------------------
 * Replaces DistillationLoss.forward + its autograd (distillation_loss.py:14-128) for the dense
 * teacher.  z = student logits, y = teacher logits, both [B,T,V] with arbitrary b/t strides.
 * n_norm: device int32 holding the normaliser N used in the gradient (this rank's N from
 * kd_prepare_rows, or the all-reduced global N).  grad_scale: host scalar folded into dlogits.
 * dlogits: NULL (forward only: one sweep, 4 B/element for bf16) or a contiguous [B,T,V] buffer of
 * z's dtype that receives d(total)/dz * grad_scale for EVERY row (zeros on invalid rows).
 * -inf entries: a -inf teacher logit has probability 0 and contributes 0 (xlogy convention of nn.KLDivLoss, :68);
 * a -inf student logit gives what the reference gives - an infinite KL where the teacher has mass on that column,
 * NaN where the teacher holds -inf there too - with CE, the teacher monitor and other rows' gradients unaffected.
 * workspace: kd_stream_workspace_bytes() bytes, 16-byte aligned. */
size_t kd_stream_workspace_bytes(void);
int kd_dense_fwd_bwd(const void* z, int z_dtype, int64_t z_stride_b, int64_t z_stride_t,
                     const void* y, int y_dtype, int64_t y_stride_b, int64_t y_stride_t,
                     const int32_t* row_target, int B, int T, int V, float tau, float alpha,
                     const int32_t* n_norm, float grad_scale, float* sums, void* dlogits,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Sparse teacher (distillation_loss.py:73-118): topk_v float32 [B,T,K] teacher log-probs at
 * tau = 1, topk_i int32 [B,T,K] vocabulary indices (both contiguous).  K <= 1024. */
int kd_sparse_fwd_bwd(const void* z, int z_dtype, int64_t z_stride_b, int64_t z_stride_t,
                      const float* topk_v, const int32_t* topk_i, int K,
                      const int32_t* row_target, int B, int T, int V, float tau, float alpha,
                      const int32_t* n_norm, float grad_scale, float* sums, void* dlogits,
                      void* workspace, size_t workspace_bytes, void* stream);

/* x[i] *= *scale for i < n, skipped on device when *scale == 1 (upstream grad_output of the
 * scalar loss; accelerate divides the loss by the accumulation steps, train.py:349). */
int kd_scale_inplace(void* x, int dtype, int64_t n, const float* scale, void* stream);

/* ---- K3: teacher top-k log-prob compaction -----------------------------------------------
 * train.py:82-91 and extract_teacher_logits.py:114-129: log_softmax over V, top-k, values to
 * fp16, indices to int32.  logits [R,V] (row stride in elements), out_v fp16 [R,k], out_i
 * int32 [R,k].  Output order: logit descending, ties by ascending index.  k <= 512. */
int kd_topk_logprobs(const void* logits, int dtype, int64_t R, int V, int64_t row_stride, int k,
                     void* out_v, int32_t* out_i, void* stream);

/* The same call with a caller-provided workspace of kd_topk_workspace_bytes(R, V) bytes (256-byte aligned; about 1/32
 * of the logits): the two-kernel form - a statistics sweep over (row, 8192-element segment) work items (piece maxima
 * + partial log-sum-exp records into the workspace) beside the selection of the previous row block, which reads the
 * ~k pieces of a row that can hold a top-k entry.  Same outputs as kd_topk_logprobs (values agree to the last bit of
 * the fp32 log-sum-exp: other summation order).  Unaligned rows, k > 128 or a null / short workspace fall back to
 * kd_topk_logprobs.  Replaces extract_teacher_logits.py:114-117 / train.py:82-91 like kd_topk_logprobs. */
size_t kd_topk_workspace_bytes(int64_t R, int V);
int kd_topk_logprobs_ws(const void* logits, int dtype, int64_t R, int V, int64_t row_stride, int k,
                        void* out_v, int32_t* out_i, void* workspace, size_t workspace_bytes, void* stream);

/* ---- token-shard gradient all-reduce through the NVSwitch (NVLS multicast) ---------------------
 * The SUM all-reduce of lm_head.weight.grad that train.py's data-parallel wrapper (accelerate / DDP) performs, for a
 * gradient that lives in a symmetric buffer with one multicast address (torch.distributed._symmetric_memory): this
 * rank reduces its 1/world share of [byte_offset, byte_offset + bytes) in the switch (multimem.ld_reduce, fp32
 * accumulation) and stores the sums into every replica (multimem.st).  The caller brackets the call with cross-GPU
 * barriers on the same stream (all replicas written before, all sums stored after).  dtype: KD_DTYPE_BF16 / _F32. */
int kd_multimem_allreduce(void* multicast_base, size_t byte_offset, size_t bytes, int dtype, int rank, int world,
                          int ctas, void* stream);

/* ---- measurement aid: read-only streaming bandwidth of this GPU ------------------------------
 * Not on the reference's path: bench.py / tools/read_probe.py time it to state what a read-only stream (K2 forward,
 * K3) can reach next to MEASURED_PEAKS.json's copy figure.  mode 0: ld.global.nc 16-byte loads, `unroll` in flight
 * per thread; mode 1: the same with an L2 evict_first policy; mode 2: cp.async.bulk (TMA) into a shared-memory ring
 * of `unroll` 16 KB stages.  scratch4: 4 writable device bytes. */
int kd_probe_read_bandwidth(const void* p, size_t bytes, int mode, int ctas_per_sm, int unroll, void* scratch4,
                            void* stream);

/* ---- stage1 frozen-vocabulary row mask -----------------------------------------------------
 * stage1.py:53-57 / 67-71: grad[:old_vocab] = 0, in place on a [V,H] gradient. */
int kd_mask_rows(void* grad, int dtype, int64_t old_vocab, int64_t H, void* stream);

/* ---- valid-row compaction (optional front end of K1) ----------------------------------------
 * distillation_loss.py:37-45 drops the rows whose label is ignored (text prefix, padding: data.py:246-251)
 * with a boolean gather; kd_compact_rows does it on the device without a host sync:
 *   perm[j]     original row of the j-th valid row (j < N, order preserved), -1 behind
 *   inv[r]      rank of row r among the valid rows, -1 if the row is not scored
 *   target_c[j] row_target[perm[j]], -1 behind            n_valid (optional) = N
 * kd_gather_rows: dst[j,:] = src[map[j],:] (rows of row_bytes bytes; strides in bytes) for map[j] >= 0, zeros
 * (zero_fill != 0) or untouched otherwise - hidden / teacher rows in with map = perm, dH back with map = inv.
 * The fused entry points below take n_rows = device pointer to N (or NULL): with compacted operands every GEMM
 * tile behind row N is skipped and the dW contraction stops at the last live row block. */
int kd_compact_rows(const int32_t* row_target, int R, int32_t* perm, int32_t* inv, int32_t* target_c,
                    int32_t* n_valid, void* stream);
int kd_gather_rows(const void* src, int64_t src_stride_bytes, const int32_t* map, int R, void* dst,
                   int64_t dst_stride_bytes, int64_t row_bytes, int zero_fill, void* stream);
/* Zero-fills dst (16-byte aligned, bytes % 16 == 0) iff *n_rows == 0, else returns at once: with no live row the
 * dW contraction is skipped and dW would stay unwritten (N == 0 must give zero gradients, :47-53). */
int kd_zero_if_empty(void* dst, int64_t bytes, const int32_t* n_rows, void* stream);

/* ---- K1: fused LM head + KD (logits never materialised) ------------------------------------
 * Replaces lm_head (transformers Qwen3ForCausalLM.lm_head, called at train.py:54) followed by
 * DistillationLoss.forward, and their backward.  h [R,H] bf16 (R = B*T rows, row stride
 * h_stride), W [V,H] bf16 (row stride w_stride), row_target from kd_prepare_rows.
 * Teacher: dense y [R,V] (y_dtype bf16/f32, row stride y_stride), sparse (topk_v fp32 [R,K] teacher
 * log-probs at tau = 1, topk_i int32 [R,K], both contiguous, K <= 1024; duplicate indices accumulate,
 * indices outside [0,V) are ignored) or none (alpha is forced to 1: plain causal-LM CE, stage1).
 * The sparse form is distillation_loss.py:73-118 without the student log-softmax over [rows,V]:
 * sum_k p_k z[i_k] is picked out of the accumulator tiles and P = scatter(i_k, p_k) is subtracted
 * from the gradient tiles in the epilogues.
 *
 * kd_fused_linear_fwd  : sums[8] and row_stats float[R,4] = (LSE1, LSEtau, LSEteacher_tau, valid).
 * kd_fused_linear_bwd  : dH [R,H] and dW [V,H] in grad_dtype (KD_DTYPE_BF16 = what autograd hands a bf16
 *                        parameter; KD_DTYPE_F32 = unrounded accumulators, for verification) (rows < dw_row_begin are NOT written:
 *                        stage1 passes old_vocab and zero-fills once, stage1.py:46-57);
 *                        grad_coef: device float[2] = (w_ce, w_kl); the gradient returned is
 *                        d[(w_ce * sum CE + w_kl * tau^2 * sum KL) / N]; the usual call passes
 *                        (alpha * g, (1 - alpha) * g) with g the upstream grad of the total loss
 *                        (distillation_loss.py:126); n_norm as above.
 * Inside the backward the gradient tile travels through the tensor cores as power-of-two-scaled fp16 against fp16
 * copies of h and of the chunk's W rows (fp32 accumulation; KD_G_FP16=0 in the environment selects bf16 operands).
 * v_chunk: vocabulary columns per backward chunk (0 = library default); the gradient scratch is
 * 2 x R x v_chunk bf16 (double buffered), independent of V.  workspace sized by
 * kd_fused_workspace_bytes() (K = top-k width of a sparse teacher, else 0), 256-byte aligned.
 * The backward runs its three chains (gradient chunk, dW, dH) on the caller's stream plus two
 * internal streams that fork from and join back into it, so the call is still ordered like one
 * stream operation; KD_BWD_STREAMS=0 in the environment keeps everything on the caller's stream.
 *
 * Logit cache (optional, NULL / 0 = off).  The forward can keep the first columns of the logits, encoded in 16 bits
 * (fp16 of z minus a per-row, per-32-column integer reference, so the rounding error of an entry shrinks with its
 * probability), in a caller-provided buffer; the backward then forms those vocabulary chunks of the gradient with an
 * HBM-bound elementwise kernel that runs beside the dW / dH GEMMs instead of recomputing the logits with a fourth
 * GEMM (executed FLOPs 8 R H V -> 6 R H V when everything fits).  The buffer size is the caller's choice, NOT a
 * function of V: kd_fused_logit_cache_bytes(R, V, v_chunk, budget) returns how much of `budget` is usable (whole
 * backward chunks: about 514 bytes per row and 256 columns); columns that do not fit are recomputed as before, so
 * peak memory stays bounded by workspace + budget for any vocabulary.  Pass the same buffer, size and v_chunk to the
 * backward of the same forward.  256-byte aligned. */
size_t kd_fused_workspace_bytes(int R, int H, int V, int v_chunk, int K);
size_t kd_fused_logit_cache_bytes(int R, int V, int v_chunk, size_t budget_bytes);
int kd_fused_linear_fwd(const void* h, int64_t h_stride, const void* W, int64_t w_stride,
                        int teacher_kind, const void* y, int y_dtype, int64_t y_stride,
                        const float* topk_v, const int32_t* topk_i, int K,
                        const int32_t* row_target, const int32_t* n_rows, int R, int H, int V, float tau,
                        float alpha, float* sums, float* row_stats, void* logit_cache, size_t logit_cache_bytes,
                        void* workspace, size_t workspace_bytes, void* stream);
int kd_fused_linear_bwd(const void* h, int64_t h_stride, const void* W, int64_t w_stride,
                        int teacher_kind, const void* y, int y_dtype, int64_t y_stride,
                        const float* topk_v, const int32_t* topk_i, int K,
                        const int32_t* row_target, const int32_t* n_rows, const float* row_stats, int R, int H,
                        int V, float tau, const int32_t* n_norm, const float* grad_coef, int grad_dtype,
                        void* dH, int64_t dh_stride, void* dW, int64_t dw_stride, int64_t dw_row_begin,
                        int v_chunk, const void* logit_cache, size_t logit_cache_bytes, void* workspace,
                        size_t workspace_bytes, void* stream);

/* The same backward restricted to vocabulary rows/columns [v_begin, v_end) (v_begin a multiple of 256), so a
 * data-parallel caller can all-reduce finished dW row blocks while later ones are computed:
 * dW rows of the range are final when the call's work completes; dH is accumulated in the workspace across
 * calls (same workspace every call): KD_RANGE_FIRST starts the accumulation, KD_RANGE_LAST writes dH.
 * dw_ready_stream (may be null): the stream that will consume a finished range's dW rows (the gradient all-reduce).
 * Non-null: after every range but the last, THAT stream waits for the rows and the caller's stream does not, so the
 * next range's kernels follow without a pipeline drain; the last range always joins into `stream`.
 * sm_limit > 0 caps the SMs the GEMM kernels occupy (leave the rest to the collective's CTAs).
 * v_offset: 0, or the first vocabulary index of this rank's slice in vocab-parallel mode (see below). */
#define KD_RANGE_FIRST 1
#define KD_RANGE_LAST 2
int kd_fused_linear_bwd_range(const void* h, int64_t h_stride, const void* W, int64_t w_stride,
                              int teacher_kind, const void* y, int y_dtype, int64_t y_stride,
                              const float* topk_v, const int32_t* topk_i, int K, const int32_t* row_target,
                              const int32_t* n_rows, const float* row_stats, int R, int H, int V, float tau,
                              const int32_t* n_norm,
                              const float* grad_coef, int grad_dtype, void* dH, int64_t dh_stride, void* dW,
                              int64_t dw_stride, int64_t dw_row_begin, int v_chunk, int v_begin, int v_end,
                              int range_flags, int sm_limit, int v_offset, const void* logit_cache,
                              size_t logit_cache_bytes, void* workspace, size_t workspace_bytes,
                              void* dw_ready_stream, void* stream);

/* Measurement hooks (no arithmetic).  kd_launch_count: kernels this library has launched in this process so far.
 * kd_fused_bwd_trace_begin arms a trace of the following kd_fused_linear_bwd* calls: every kernel they launch is
 * bracketed by two timed CUDA events on the stream it is launched on (the backward runs on three streams);
 * kd_fused_bwd_trace_read synchronises the device, disarms the trace and writes up to max_records HOST records
 * float[4] = (class, vocabulary chunk, start ms, end ms) relative to the first event; classes: 0 fp16 operand copy,
 * 1 gradient chunk from the logit cache, 2 dW GEMM, 3 dH GEMM, 4 gradient chunk by recompute GEMM.  Returns the
 * number of records, -1 on error. */
unsigned long long kd_launch_count(void);
int kd_fused_bwd_trace_begin(void);
int kd_fused_bwd_trace_read(float* host_out, int max_records);

/* ---- vocab-parallel mode: every rank holds W[v_offset : v_offset + V, :] (V = its slice) and the
 * matching teacher columns, all ranks see the same rows.  kd_fused_linear_fwd_partial runs the forward over the
 * slice and writes one 12-float record per row (running max / sums of the 4 soft-maxes, KL cross term, label
 * logits owned by the slice, sparse row constants); the caller all-gathers the records [G][R][12] and
 * kd_fused_merge_ranks applies the split-V merge rule (appendix C) -> sums[8] + row_stats[R,4], identical on every
 * rank.  The backward is kd_fused_linear_bwd_range on the slice with the same v_offset (labels and top-k indices
 * stay global): dW rows of the slice are final, dH is this slice's partial sum (all-reduce it in fp32). */
int kd_fused_linear_fwd_partial(const void* h, int64_t h_stride, const void* W, int64_t w_stride,
                                int teacher_kind, const void* y, int y_dtype, int64_t y_stride,
                                const float* topk_v, const int32_t* topk_i, int K,
                                const int32_t* row_target, const int32_t* n_rows, int R, int H, int V, int v_offset,
                                float tau, float* rank_rec, void* logit_cache, size_t logit_cache_bytes,
                                void* workspace, size_t workspace_bytes, void* stream);
size_t kd_fused_merge_workspace_bytes(void);
int kd_fused_merge_ranks(const float* rank_recs, int G, const int32_t* row_target, int R, int teacher_kind,
                         float tau, float* sums, float* row_stats, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ---- stage-1 names: causal-LM cross-entropy through the LM head, no teacher --------
 * What TRL's SFTTrainer asks of the model in stage1.py:329-340 (transformers ForCausalLMLoss, or Liger's
 * fused-linear-CE under use_liger_kernel, :315), with stage1.py:46-57's frozen-vocabulary mask folded into the
 * dW GEMM: pass dw_row_begin = V - num_new_tokens and rows below it are never computed (zero-fill dW once).
 * Thin forwards of kd_fused_linear_fwd / _bwd with KD_TEACHER_NONE, tau = alpha = 1; grad_coef[0] is the
 * upstream gradient of the mean CE. */
int kd_ce_fused_linear_fwd(const void* h, int64_t h_stride, const void* W, int64_t w_stride,
                           const int32_t* row_target, const int32_t* n_rows, int R, int H, int V,
                           float* sums, float* row_stats, void* logit_cache, size_t logit_cache_bytes,
                           void* workspace, size_t workspace_bytes, void* stream);
int kd_ce_fused_linear_bwd(const void* h, int64_t h_stride, const void* W, int64_t w_stride,
                           const int32_t* row_target, const int32_t* n_rows, const float* row_stats, int R,
                           int H, int V, const int32_t* n_norm, const float* grad_coef, int grad_dtype,
                           void* dH, int64_t dh_stride, void* dW, int64_t dw_stride, int64_t dw_row_begin,
                           int v_chunk, const void* logit_cache, size_t logit_cache_bytes, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ---- teacher LM head in front of the top-k compaction ---------------------------------------
 * out[R,V] bf16 (row stride out_stride, a multiple of 8 elements, 16-byte aligned base; the padding columns
 * V .. roundup(V, 8) - 1 of a row may be zero-filled: TMA stores whole 16-byte granules) = h[R,H] * W[V,H]^T,
 * fp32 accumulation, one rounding to bf16 - what transformers' bf16 lm_head produces for teacher_outputs.logits
 * (train.py:60-72, extract_teacher_logits.py:109-113).  The host mirror runs it over blocks of rows into a
 * fixed-size scratch and feeds kd_topk_logprobs, so the teacher's [B,T,V] logits never exist as a whole. */
int kd_linear_bf16(const void* h, int64_t h_stride, const void* W, int64_t w_stride, void* out,
                   int64_t out_stride, int R, int H, int V, void* stream);

/* ---- teacher LM head fused with the top-k selection statistics (train.py:60-94 on-the-fly top-k,
 * extract_teacher_logits.py:109-129) -----------------------------------------------------------------------------
 * kd_head_logits_stats = kd_linear_bf16 whose epilogue also leaves, computed from the same bf16-rounded logits,
 *   pmax [R][pmax_stride] bf16 : maximum of every 32-column piece of a row (pieces past V hold -inf),
 *   part [R][part_stride] float2: partial (max, sum exp(x - max)) records of the row; *n_part (host int) receives how
 *                                 many of them are written per row for this (R, V);
 * kd_head_topk_select then produces log_softmax -> top-k (fp16 values, int32 indices; same ordering, tie rule and
 * value rounding as kd_topk_logprobs on these logits) by merging the records and reading only the pieces whose
 * maximum reaches the k-th largest one - a few KB per row instead of the row.  Indices are identical to
 * kd_topk_logprobs on the same scratch; values agree to the last bit of the fp32 log-sum-exp (other summation order).
 * kd_head_topk_layout reports the strides to allocate for a vocabulary of V columns (any R). */
int kd_head_topk_layout(int V, int* pmax_stride, int* part_stride);
int kd_head_logits_stats(const void* h, int64_t h_stride, const void* W, int64_t w_stride, void* out,
                         int64_t out_stride, void* pmax, int pmax_stride, void* part, int part_stride,
                         int* n_part, int R, int H, int V, void* stream);
int kd_head_topk_select(const void* logits, int64_t row_stride, const void* pmax, int pmax_stride,
                        const void* part, int part_stride, int n_part, int64_t R, int V, int k, void* out_v,
                        int32_t* out_i, void* stream);

/* Plain bf16 GEMM on the same tcgen05 pipeline (test hook for the K1 building block), fp32 out:
 *   C[M,N] (ldc) = op(A) * op(B)^T with
 *   a_mn_major = 0: A is [M][K] (K contiguous, lda)   | 1: A is [K][M] (M contiguous, lda)
 *   b_mn_major = 0: B is [N][K] (K contiguous, ldb)   | 1: B is [K][N] (N contiguous, ldb)
 * (a_mn_major = 1, b_mn_major = 0) is not instantiated. */
int kd_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
                 float* C, int64_t ldc, int M, int N, int K, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KD_B200_H_ */
