/* glnn_b200.h -- C ABI of libglnn_b200.so, the B200 (sm_100a) implementation of the GLNN hot path.
 *
 * The reference (snap-research/graphless-neural-networks @ 80d0f8ea) is pure Python and has no FFI;
 * its extension points are the `models.Model` class, the step functions of train_and_eval.py and
 * the out.npz file (README.md:77-79).  Each entry point below names the
 * reference call it replaces.  The Python side (glnn_b200/_lib.py) binds them with ctypes; the stub
 * a maintainer would add to the reference is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is CALLER-OWNED DEVICE memory unless the name ends in `_host`; the library never
 *     frees caller memory and keeps no reference to it after the call returns;
 *   - matrices are row-major fp32 with an explicit leading dimension (elements);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*) and is asynchronous; no entry
 *     point synchronises the device unless documented;
 *   - return value: 0 = success; < 0 = argument/shape/alignment error, nothing was launched;
 *     > 0 = a cudaError_t passed through.  glnn_last_error() returns a thread-local message.
 */
#ifndef GLNN_B200_H_
#define GLNN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLNN_ABI_VERSION 2

#define GLNN_ERR_ARG (-1)       /* null pointer / negative size / unsupported flag combination  */
#define GLNN_ERR_SHAPE (-2)     /* dimension outside the supported range                        */
#define GLNN_ERR_ALIGN (-3)     /* pointer or leading dimension violates an alignment rule      */
#define GLNN_ERR_WORKSPACE (-4) /* workspace too small (see the matching *_workspace_bytes)     */
#define GLNN_ERR_DEVICE (-5)    /* not an sm_100 device                                         */

#if defined(__GNUC__)
#define GLNN_API __attribute__((visibility("default")))
#else
#define GLNN_API
#endif

typedef void* glnn_stream_t;

GLNN_API int glnn_version(void);
GLNN_API const char* glnn_last_error(void);
/* Fills SM count and compute capability of the current device.  Fails with GLNN_ERR_DEVICE when the
 * device is not compute capability 10.x. */
GLNN_API int glnn_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * K1/K2/K3: CSR neighbour aggregation  (dgl update_all(copy_u, sum) + the scalings around it)
 *
 *   acc[v,:] = sum_{e in [indptr[v], indptr[v+1])} src_scale[indices[e]] * X[indices[e], :]
 *   if self_add:        acc[v,:] += X[v,:]                  (dst nodes are a prefix of src nodes)
 *   if mean_plus_one:   acc[v,:] /= (in_deg(v) + 1)         (SAGEConv "gcn", models.py:112,138)
 *   if dst_scale:       acc[v,:] *= dst_scale[v]            (GraphConv norm="both", models.py:193)
 *   t      = acc[v,c] + bias[c];            if relu == 2: t = max(t, 0)   (GCN: relu -> norm)
 *   t      = t * col_scale[c] + col_shift[c];  if relu == 1: t = max(t, 0)   (SAGE: norm -> relu)
 *   Y[v,c] = t
 *
 * Replaces: dgl.nn.SAGEConv("gcn") aggregation (models.py:84-99,112,138), dgl.nn.GraphConv
 * aggregation (models.py:170-187,193) and g.update_all(copy_u,sum) in utils.py:185.
 * indptr is int32 or int64 (indptr64), indices int32.  src_scale, dst_scale, bias, col_scale,
 * col_shift may be NULL.  Fast path needs d % 4 == 0, ldx % 4 == 0, ldy % 4 == 0 and 16-byte
 * aligned X/Y; anything else takes a scalar path.  Multi-edges count with multiplicity; an empty
 * row yields the epilogue applied to (self_add ? X[v] : 0).
 */
GLNN_API int glnn_spmm_csr_f32(const void* indptr, int indptr64, const int32_t* indices, const float* X,
                      int64_t ldx, float* Y, int64_t ldy, int64_t n_dst, int64_t n_src, int d,
                      int self_add, int mean_plus_one, const float* src_scale,
                      const float* dst_scale, const float* bias, const float* col_scale,
                      const float* col_shift, int relu, glnn_stream_t stream);

/* Same aggregation, result written as bf16 hi / lo planes (see "tensor-core operand format" below)
 * for a projection that follows: Y_hi / Y_lo are [n_dst, ldyp] uint16, ldyp % 4 == 0. */
GLNN_API int glnn_spmm_csr_planes(const void* indptr, int indptr64, const int32_t* indices, const float* X,
                         int64_t ldx, uint16_t* Y_hi, uint16_t* Y_lo, int64_t ldyp, int64_t n_dst,
                         int64_t n_src, int d, int self_add, int mean_plus_one,
                         const float* src_scale, const float* dst_scale, const float* bias,
                         const float* col_scale, const float* col_shift, int relu,
                         glnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K4: fp32-faithful dense projection with fused epilogue
 *
 *   t      = row_scale[m] * sum_k opA(m,k) * opB(k,n) + bias[n];   if relu == 2: t = max(t, 0)
 *   C[m,n] = t * col_scale[n] + col_shift[n];                      if relu == 1: C = max(C, 0)
 *   opA(m,k) = transA ? A[k*lda + m] : A[m*lda + k]
 *   opB(k,n) = transB ? B[n*ldb + k] : B[k*ldb + n]      (transB=1 is an nn.Linear weight [N,K])
 *
 * Replaces: nn.Linear inside SAGEConv.fc_neigh / MLP.layers (models.py:46,84-99), torch.matmul in
 * GraphConv (models.py:170-187), and their autograd backward GEMMs in train_mini_batch
 * (train_and_eval.py:82-85).  Accumulation is fp32 (or error-compensated 3xTF32 on the tensor
 * path, which holds the 1e-4 parity bound).  Any epilogue pointer may be NULL.
 * `impl`: 0 = auto, 1 = force SIMT fp32, 2 = force tcgen05 3xTF32 (error if the shape is not
 * eligible).
 */
GLNN_API int glnn_gemm_f32(const float* A, int64_t lda, int transA, const float* B, int64_t ldb, int transB,
                  float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, const float* row_scale,
                  const float* bias, const float* col_scale, const float* col_shift, int relu,
                  int impl, glnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K4, tensor-core operand format.  A "planes" matrix is an fp32 matrix kept as two bf16 matrices
 * hi = bf16(x), lo = bf16(x - hi) (uint16_t storage) with a common leading dimension ldp that is
 * a multiple of 8 elements and zeros in the pad columns; x ~ hi + lo to 2^-18 relative, same
 * bytes as fp32.  glnn_split_planes_f32 converts; glnn_gemm_bf16x3_planes computes the same
 * C = epilogue(op(A) op(B)) as glnn_gemm_f32 from planes (tcgen05, hi*hi + hi*lo + lo*hi in fp32)
 * and writes C as fp32 (C != NULL) and/or as planes (C_hi/C_lo != NULL, for a following GEMM).
 * Kernels of this library that only feed GEMMs write planes directly.
 */
GLNN_API int glnn_split_planes_f32(const float* X, int64_t ldx, int64_t rows, int cols, uint16_t* hi,
                          uint16_t* lo, int64_t ldp, glnn_stream_t stream);

GLNN_API int glnn_gemm_bf16x3_planes(const uint16_t* A_hi, const uint16_t* A_lo, int64_t lda, int transA,
                            const uint16_t* B_hi, const uint16_t* B_lo, int64_t ldb, int transB,
                            float* C, int64_t ldc, uint16_t* C_hi, uint16_t* C_lo, int64_t ldcp,
                            int64_t M, int64_t N, int64_t K, const float* row_scale,
                            const float* bias, const float* col_scale, const float* col_shift,
                            int relu, glnn_stream_t stream);

/* 24-bit row-packed matrices ("q24"): a matrix that is only read by a neighbour GATHER is stored per
 * row as dq hi16 values followed by dq mid8 values (dq = d rounded up to 8; the top 24 bits of each
 * fp32, rounded to nearest: 2^-17 relative; pad columns zero), rows ldq bytes apart with
 * ldq >= 3 dq and ldq % 16 == 0 -- glnn_q24_row_bytes(d) rounds 3 dq up to whole 32-byte DRAM
 * sectors.  The gather is DRAM-bound on bytes per neighbour row, so this is 25 % less traffic on
 * the dominant kernel of the teacher forward.
 * glnn_quantize_q24_f32 converts an fp32 matrix (the caller's features);
 * glnn_gemm_bf16x3_planes_q24 = glnn_gemm_bf16x3_planes (transA = 0) with the q24 output only
 * (N % 8 == 0); glnn_spmm_csr (below) reads q24 through desc->X_q24. */
GLNN_API int64_t glnn_q24_row_bytes(int d);

GLNN_API int glnn_quantize_q24_f32(const float* X, int64_t ldx, int64_t rows, int d, uint8_t* X_q24,
                          int64_t ldq, glnn_stream_t stream);

GLNN_API int glnn_gemm_bf16x3_planes_q24(const uint16_t* A_hi, const uint16_t* A_lo, int64_t lda,
                                const uint16_t* B_hi, const uint16_t* B_lo, int64_t ldb, int transB,
                                uint8_t* C_q24, int64_t ldq, int64_t M, int64_t N, int64_t K,
                                const float* row_scale, const float* bias, const float* col_scale,
                                const float* col_shift, int relu, glnn_stream_t stream);

/* Legacy form: q24 rows of exactly 3 d bytes (d % 16 == 0), planes output. */
GLNN_API int glnn_spmm_csr_q24_planes(const void* indptr, int indptr64, const int32_t* indices,
                             const uint8_t* X_q24, uint16_t* Y_hi, uint16_t* Y_lo, int64_t ldyp,
                             int64_t n_dst, int64_t n_src, int d, int self_add, int mean_plus_one,
                             const float* src_scale, const float* dst_scale, glnn_stream_t stream);

/* General form of the aggregation (same math as glnn_spmm_csr_f32): input fp32 (X, ldx) or q24
 * (X_q24, ldq); output fp32 (Y, ldy) and/or bf16 planes (Y_hi, Y_lo, ldyp); with a q24 input the
 * output rows are written in whole 8-column chunks (ldy / ldyp >= d rounded up to 8, pad = 0).
 *   log_softmax = c > 0: the epilogue ends with log_softmax over the first c columns and writes only
 *     those (Y fp32, any ldy >= c): evaluate()'s log_softmax (train_and_eval.py:98) fused into the
 *     last layer's aggregation.  d <= 512.
 *   Y_init != NULL: the row sums start from Y_init[v, :] (before the self term / mean / scales):
 *     acc = Y_init + sum over the edges.  Lets a caller split the edge set of an aggregation into
 *     passes (first pass: fp32 Y, no epilogue; last pass: Y_init = that Y, full epilogue).
 *   hot_below = k > 0: L2 residency hint.  Gathered rows of source ids < k are loaded with the
 *     evict_last policy, every other access of the kernel (cold rows, indices, output) evict_first,
 *     so that the most-referenced rows of a degree-ordered graph stay in the 126 MB L2.  Purely a
 *     performance hint: results do not depend on it.
 */
typedef struct glnn_spmm_desc {
  const void* indptr;     /* int32 or int64 [n_dst + 1] */
  const int32_t* indices; /* [nnz] */
  int32_t indptr64;
  int32_t d;
  int64_t n_dst, n_src;
  const float* X;
  int64_t ldx;
  const uint8_t* X_q24;
  int64_t ldq;
  float* Y;
  int64_t ldy;
  uint16_t* Y_hi;
  uint16_t* Y_lo;
  int64_t ldyp;
  int32_t self_add, mean_plus_one;
  const float* src_scale;
  const float* dst_scale;
  const float* bias;
  const float* col_scale;
  const float* col_shift;
  int32_t relu;
  int32_t log_softmax;
  int32_t hot_below;
  int32_t reserved;
  const float* Y_init;    /* optional: acc[v, :] starts from Y_init[v, 0:d] (fp32) instead of 0 -- the   */
  int64_t ldyi;           /* second pass of an aggregation split by source block (sharded teacher)   */
} glnn_spmm_desc;

GLNN_API int glnn_spmm_csr(const glnn_spmm_desc* desc, glnn_stream_t stream);

/* EXPERIMENTAL, opt-in (nothing on the default path calls it; added at the end of round 1 without a
 * GPU run -- DESIGN.md section 8 item 3): lossless sparse rows for post-ReLU embeddings.  "s24" row =
 * the row's non-zeros as 32-bit words [fp32 bits 31..8 | column 7..0] sorted by column, followed by
 * zero words up to lds (a multiple of 32 words, >= d rounded up to 32 = glnn_s24_row_words(d));
 * d <= 256.  glnn_compact_s24 derives it from the q24 matrix and leaves the largest non-zero count
 * of a row in *cap_dev (device memory).  glnn_spmm_csr_s24 = glnn_spmm_csr on desc (which must carry
 * the q24 matrix: self term, hub rows and the dense fallback read it) gathering non-hub rows from the
 * sparse copy whenever cap * 4 bytes <= 80 % of the q24 row; results are identical to glnn_spmm_csr
 * except for the summation order inside hub rows.  128 < d <= 256, no src_scale, no log_softmax. */
GLNN_API int64_t glnn_s24_row_words(int d);
GLNN_API int glnn_compact_s24(const uint8_t* X_q24, int64_t ldq, int64_t rows, int d, uint32_t* X_s24,
                     int64_t lds, int32_t* cap_dev, glnn_stream_t stream);
GLNN_API int glnn_spmm_csr_s24(const glnn_spmm_desc* desc, const uint32_t* X_s24, int64_t lds,
                      const int32_t* cap_dev, glnn_stream_t stream);

/* EXPERIMENT, measurement only (tools/exp_spmm_tma.py; nothing on the product path calls it): the
 * aggregation of 256-wide q24 rows (ldq = 768) with every neighbour row pulled by ONE TMA bulk copy
 * (cp.async.bulk global -> shared, mbarrier complete_tx) into a per-warp ring of `stages` (4 or 8)
 * row buffers instead of ld.global.nc into registers -- the A/B that settles the north_star's "TMA
 * for the neighbour pull" clause with a number.  Same math as glnn_spmm_csr; no hub path. */
GLNN_API int glnn_exp_spmm_tma_q24(const glnn_spmm_desc* desc, int stages, glnn_stream_t stream);

/* K5 (eval): folds BatchNorm1d running statistics into a per-column affine for the epilogues
 * above: scale = gamma / sqrt(var + eps), shift = beta - mean * scale  (models.py:139-141). */
GLNN_API int glnn_bn_fold_f32(const float* gamma, const float* beta, const float* mean, const float* var,
                     float eps, float* scale, float* shift, int n, glnn_stream_t stream);

/* K7: Y[r,:] = log_softmax(X[r, 0:c])  (train_and_eval.py:98,124).  In-place allowed. */
GLNN_API int glnn_log_softmax_f32(const float* X, int64_t ldx, float* Y, int64_t ldy, int64_t n, int c,
                         glnn_stream_t stream);

/* K7+K8 (eval): for the rows listed in idx (int64, NULL = all n rows) of log-probabilities LP
 * accumulates  out[0] += sum_i -LP[idx_i, labels[idx_i]]  and  out[1] += #(argmax == label)
 * (criterion + evaluator of train_and_eval.py:99-104,129-134).  `out` is 2 floats, zeroed by
 * the caller. */
GLNN_API int glnn_nll_acc_f32(const float* LP, int64_t ld, int c, const int64_t* labels, const int64_t* idx,
                     int64_t n, float* out, glnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Student (B): MLP distillation.  Replaces the body of train_mini_batch (train_and_eval.py:59-86)
 * and evaluate_mini_batch (:108-136) for models.MLP (models.py:7-53) with norm_type in
 * {"none","batch"}.
 *
 * Parameter memory is FLAT: one fp32 buffer per role (params, grads, exp_avg, exp_avg_sq) with the
 * layout   [W_0 (out x in) | b_0 | W_1 | b_1 | ... | gamma_0 | beta_0 | gamma_1 | ...]   and a flat
 * buffer bn_stats = [running_mean_0 | running_var_0 | running_mean_1 | ...].  The host side keeps
 * nn.Parameter / optimizer-state tensors as views into these buffers so state_dict() and
 * torch.optim.Adam stay coherent.
 */
typedef struct glnn_mlp_desc {
  int32_t num_layers;  /* >= 1 */
  int32_t feat_dim;
  int32_t hidden_dim;
  int32_t label_dim;
  int32_t norm;        /* 0 = none, 1 = BatchNorm1d */
  float dropout;       /* drop probability p, 0 <= p < 1 */
  float bn_eps;        /* 1e-5 */
  float bn_momentum;   /* 0.1 */
} glnn_mlp_desc;

/* torch.optim.Adam hyper-parameters (L2 decay, train_student.py:275), in DOUBLE like the Python
 * floats torch reads them from: 1 - beta and the bias corrections 1 - beta^t are formed in double and
 * rounded to fp32 once, exactly as torch does (ABI version 2; version 1 carried floats). */
typedef struct glnn_adam_hparams {
  double lr, beta1, beta2, eps, weight_decay;
} glnn_adam_hparams;

/* Number of fp32 elements in the flat parameter buffer / in bn_stats. */
GLNN_API int64_t glnn_mlp_param_count(const glnn_mlp_desc* desc);
GLNN_API int64_t glnn_mlp_bn_stat_count(const glnn_mlp_desc* desc);
/* Bytes of scratch needed for batches of up to `rows` rows (train or eval). */
GLNN_API int64_t glnn_mlp_workspace_bytes(const glnn_mlp_desc* desc, int64_t rows);

/* One pass of train_mini_batch: nb optimizer steps over batches perm[i*bs : (i+1)*bs] of rows of
 * X [n, feat_dim].  target_kind 0: target = int64 labels [n] + NLLLoss (mean);  1: target = fp32
 * teacher LOG-probabilities [n, label_dim] + KLDivLoss(batchmean, log_target=True).
 * Each step: gather -> forward (BN batch statistics, running-stat update, dropout) -> log-softmax
 * -> loss -> backward of lamb*loss -> Adam step (adam_step0 + i + 1).  loss_sum[0] accumulates the
 * UNSCALED per-step mean losses (the caller zeroes it and divides by nb).
 * drop_masks: NULL -> device counter-based RNG keyed by (seed, step, layer, element); otherwise
 * uint8 keep-masks laid out [nb][num_layers-1][bs][hidden_dim] (parity mode).
 * num_batches_tracked (int64 [num_layers-1], may be NULL) is advanced by nb on the device.
 * At most 2^22 rows (nb * bs) per call; the host splits longer passes and carries adam_step0.
 * The kernel sequence of a step is captured once into a CUDA graph per (shape, buffer set) and
 * replayed nb times (set GLNN_NO_GRAPH=1 to launch kernels directly). */
GLNN_API int glnn_mlp_train_pass(const glnn_mlp_desc* desc, float* params, float* grads, float* exp_avg,
                        float* exp_avg_sq, float* bn_stats, int64_t* num_batches_tracked,
                        int64_t adam_step0, const glnn_adam_hparams* hp, const float* X, int64_t ldx,
                        const void* target, int target_kind, const int64_t* perm, int64_t nb,
                        int64_t bs, const uint8_t* drop_masks, uint64_t seed, float lamb,
                        float* loss_sum, void* workspace, int64_t workspace_bytes,
                        glnn_stream_t stream);

/* K10 alone: one torch.optim.Adam update (amsgrad = False, L2 weight decay folded into the gradient,
 * train_student.py:275-277 / train_teacher.py:234-236) of n contiguous fp32 parameters from given
 * gradients -- the same kernel the fused student step ends with.  `step` is the 1-based number of
 * this update (bias corrections 1 - beta^step are evaluated in double).  Used by teacher training
 * (train / train_sage, train_and_eval.py:12-56) and by the injected-gradient parity test. */
GLNN_API int glnn_adam_step_f32(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                       int64_t n, int64_t step, const glnn_adam_hparams* hp, glnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Student, data parallel over the GPUs of one box.  Same step as
 * glnn_mlp_train_pass on the SAME global batches (perm holds nb * bs_global rows; rank r processes
 * rows [r, r+1) * bs_global / world of every batch), same BatchNorm semantics (statistics of the
 * global batch) and the same Adam update, so the result equals the single-GPU pass up to fp32
 * summation order.  Every exchange is fused into the producing / consuming kernel over
 * peer-mapped memory (no NCCL on the step path, one CUDA graph per step):
 *   BN statistics (forward and backward) are written by the statistics kernels straight into
 *   every peer's buffer and awaited by the apply kernels; the optimizer kernel of rank r sums slice
 *   r of the gradients over all peers, applies Adam and stores the new parameters to every peer.
 *
 * Memory contract.  Every rank allocates ONE symmetric region of the same size (cuMem / CUDA IPC /
 * torch symmetric memory) and maps all of them: base[r] is rank r's region in THIS process.
 *   [0, glnn_mlp_dp_control_bytes)   control area, initialised by glnn_mlp_dp_init on every rank
 *                                    followed by a barrier across the ranks (caller's job);
 *   params, grads                    glnn_mlp_dp_flat_count elements each (the flat layout padded
 *                                    to equal 16-byte-aligned slices), anywhere behind the control
 *                                    area, at the SAME offsets on every rank, pad elements zero.
 * exp_avg / exp_avg_sq are ordinary local buffers of glnn_mlp_dp_flat_count elements; after a pass
 * only slice `rank` of them is current (slice = flat_count / world): gather the slices when the
 * optimizer state is read.  loss_sum receives this rank's share: sum over ranks = the pass's sum.
 * All ranks must call with the same arguments (nb, bs_global, seed, hp ...) in the same order.
 */
typedef struct glnn_dp_group {
  int32_t world, rank;  /* world <= 8 */
  void* base[8];
  int64_t bytes;        /* size of each region */
} glnn_dp_group;

GLNN_API int64_t glnn_mlp_dp_control_bytes(const glnn_mlp_desc* desc, int64_t bs_global, int world);
GLNN_API int64_t glnn_mlp_dp_flat_count(const glnn_mlp_desc* desc, int world);
GLNN_API int glnn_mlp_dp_init(void* region_local, int64_t control_bytes, glnn_stream_t stream);
GLNN_API int glnn_mlp_train_pass_dp(const glnn_dp_group* grp, const glnn_mlp_desc* desc, float* params,
                           float* grads, float* exp_avg, float* exp_avg_sq, float* bn_stats,
                           int64_t* num_batches_tracked, int64_t adam_step0,
                           const glnn_adam_hparams* hp, const float* X, int64_t ldx,
                           const void* target, int target_kind, const int64_t* perm, int64_t nb,
                           int64_t bs_global, const uint8_t* drop_masks, uint64_t seed, float lamb,
                           float* loss_sum, void* workspace, int64_t workspace_bytes,
                           glnn_stream_t stream);

/* Eval-mode forward of n contiguous rows -> out [n, label_dim] (ldo): log-probabilities if
 * log_softmax != 0 (evaluate_mini_batch), raw logits otherwise (Model.forward).  Row results are
 * independent of how rows are batched, so `rows_per_chunk` only bounds scratch
 * (glnn_mlp_workspace_bytes(desc, rows_per_chunk)). */
GLNN_API int glnn_mlp_eval(const glnn_mlp_desc* desc, const float* params, const float* bn_stats,
                  const float* X, int64_t ldx, int64_t n, float* out, int64_t ldo, int log_softmax,
                  int64_t rows_per_chunk, void* workspace, int64_t workspace_bytes,
                  glnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Teacher (A): whole eval-mode forward, layer planning included.
 *
 * glnn_sage_forward  replaces SAGE.inference (models.py:121-148; full-neighbour, eval mode, where
 *   the reference's per-batch loop equals one full-graph pass per layer -- SURVEY.md section 3.2):
 *   per layer  h = fc_neigh((sum_{u->v} h_u + h_v) / (deg_v + 1));  hidden layers: BN(eval) -> ReLU.
 *   A layer whose (4-padded) d_out is smaller than d_in is projected first (legal by linearity;
 *   it only changes fp32 rounding), otherwise aggregated first as DGL 0.6.1 does.
 * glnn_gcn_forward  replaces GCN.forward in eval mode (models.py:189-199) over GraphConv
 *   norm="both": h = act(D_in^-1/2 A (D_out^-1/2 h) W + b), W first iff d_in > d_out (DGL's rule),
 *   ReLU inside the conv then BN(eval) on hidden layers.  src_norm/dst_norm are the clamped
 *   out/in-degree^-1/2 vectors [n].
 * The `log_softmax` argument is a flag word: GLNN_FWD_LOG_SOFTMAX (1) makes the output
 * log-probabilities (evaluate, train_and_eval.py:97-98); GLNN_FWD_EXACT (2) selects plain fp32
 * arithmetic end to end (fp32 gathers, SIMT fp32 projections, DGL's operation order): several times
 * slower, but its error is fp32 rounding only (~1e-6 of max|logit|), so the raw LOGITS meet
 * allclose(rtol=1e-4, atol=1e-5) as well; the default mode holds max|a-b|/max|b| <= 1e-4 on logits
 * (measured ~1e-5) and the allclose form on the log-probabilities.
 * All pointers (including those inside glnn_gnn_layer) are DEVICE pointers.
 */
#define GLNN_FWD_LOG_SOFTMAX 1
#define GLNN_FWD_EXACT 2
typedef struct glnn_gnn_layer {
  const float* weight;   /* SAGE: [d_out, d_in] (nn.Linear);  GCN: [d_in, d_out] (GraphConv) */
  const float* bias;     /* [d_out] */
  const float* bn_scale; /* eval BN folded by glnn_bn_fold_f32; NULL when the layer has no norm */
  const float* bn_shift;
  int32_t d_in, d_out;
} glnn_gnn_layer;

GLNN_API int64_t glnn_gnn_forward_workspace_bytes(int64_t n, const glnn_gnn_layer* layers, int num_layers);

GLNN_API int glnn_sage_forward(const void* indptr, int indptr64, const int32_t* indices, int64_t n,
                      const float* X, int64_t ldx, const glnn_gnn_layer* layers, int num_layers,
                      float* out, int64_t ldo, int log_softmax, void* workspace,
                      int64_t workspace_bytes, glnn_stream_t stream);

GLNN_API int glnn_gcn_forward(const void* indptr, int indptr64, const int32_t* indices, int64_t n,
                     const float* src_norm, const float* dst_norm, const float* X, int64_t ldx,
                     const glnn_gnn_layer* layers, int num_layers, float* out, int64_t ldo,
                     int log_softmax, void* workspace, int64_t workspace_bytes,
                     glnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Sharded teacher: the per-layer embedding exchange.  glnn_peer_push copies
 * `bytes` (a multiple of 16) from this rank's slab `src` to the same data at dst[0..n_dst) -- peer-
 * mapped addresses of the other ranks' replicas (torch symmetric memory / cuMem / CUDA IPC) -- with a
 * small SM-driven kernel of `ctas` CTAs (<= 0: 32): one load, n_dst posted NVLink stores.  It replaces
 * the reference-side idea of an NCCL all-gather of layer embeddings (BASELINE.json north_star) with
 * direct pushes into the replicas the next aggregation reads; completion is ordered by the caller
 * (stream order + one device-side barrier across ranks per layer).  n_dst <= 8. */
GLNN_API int glnn_peer_push(const void* src, void* const* dst, int n_dst, int64_t bytes, int ctas,
                   glnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Teacher TRAINING: the pieces `train` (train_and_eval.py:12-29) and
 * `train_sage` (:32-56) need around the aggregation / projection kernels so that one training step is
 * a kernel sequence with a hand-written backward (no autograd); the sequence itself is laid out by
 * glnn_b200/teacher_train.py.  The backward of an aggregation is the aggregation over the transposed
 * CSR (full-batch GCN: built once) or glnn_spmm_csr_scatter_f32 (sampled blocks: no transpose).
 *
 * glnn_nll_loss_grad_f32: log_softmax + NLLLoss(mean) over m selected rows and the gradient of
 *   lamb * loss w.r.t. the logits (train_and_eval.py:21-27,46-53).  Row i of the selection is logits
 *   row r = rows ? rows[i] : i with label labels[label_rows ? label_rows[i] : r].  Writes
 *   dlogits[r, 0:c] = lamb / m * (softmax - onehot) for the selected rows only (the caller zeroes
 *   dlogits when the selection is a subset) and ADDS the unscaled mean loss to *loss_out. */
GLNN_API int glnn_nll_loss_grad_f32(const float* logits, int64_t ld, int c, const int64_t* labels,
                           const int64_t* rows, const int64_t* label_rows, int64_t m, float lamb,
                           float* dlogits, int64_t lddl, float* loss_out, glnn_stream_t stream);

/* The block between two convolutions in TRAIN mode: [BatchNorm1d] -> [ReLU] -> [Dropout]
 * (SAGE.forward, models.py:113-118) or, for GCN whose ReLU lives inside the conv, [BatchNorm1d] ->
 * [Dropout] applied to the post-ReLU conv output (models.py:194-198; relu_input = 1 makes the backward
 * end with the mask X > 0).  BatchNorm uses batch statistics, updates running_mean / running_var with
 * `momentum` (unbiased variance) and leaves mean / 1/sqrt(var + eps) in save_mean / save_invstd for
 * the backward.  Dropout keeps an element iff keep_mask[row * d + col] != 0 (parity mode) or, without
 * a mask, iff a counter-based hash of (seed, row, col) is >= p_drop; kept elements are scaled by
 * 1 / (1 - p_drop).  The backward recomputes everything from X: only X, the saved statistics and the
 * seed / mask are kept between the two calls.  scratch: 2 * d floats (BatchNorm only). */
typedef struct glnn_act_desc {
  int64_t n;                 /* rows */
  int32_t d;                 /* columns */
  int32_t relu_post;         /* SAGE: ReLU after the norm */
  const float* X;            /* [n, d] block input */
  int64_t ldx;
  float* Y;                  /* [n, d] block output (forward only) */
  int64_t ldy;
  const float* gamma;        /* BatchNorm1d weight / bias, NULL = no norm */
  const float* beta;
  float* running_mean;       /* updated by the forward; may be NULL */
  float* running_var;
  float* save_mean;          /* [d] */
  float* save_invstd;        /* [d] */
  float eps, momentum;
  float p_drop;
  int32_t relu_input;        /* GCN: X = relu(z); the backward multiplies by (X > 0) */
  uint64_t seed;
  const uint8_t* keep_mask;  /* optional [n, d] */
} glnn_act_desc;

GLNN_API int glnn_act_train_fwd_f32(const glnn_act_desc* desc, float* scratch, glnn_stream_t stream);
/* dX = gradient w.r.t. X given dY; dgamma / dbeta (BatchNorm, may be NULL); dbias (may be NULL) =
 * column sums of dX = gradient of a bias added in front of the block. */
GLNN_API int glnn_act_train_bwd_f32(const glnn_act_desc* desc, const float* dY, int64_t lddy, float* dX,
                           int64_t lddx, float* dgamma, float* dbeta, float* dbias, float* scratch,
                           glnn_stream_t stream);

/* Transposed aggregation by scatter: dX[u, :] += scale[v] * dY[v, :] for every edge u -> v of the CSR
 * over destinations (+ dX[v, :] += scale[v] * dY[v, :] with self_add): the backward of
 * glnn_spmm_csr_f32(self_add, dst_scale = scale) for a sampled block, without building its transpose.
 * dX must be initialised by the caller (zeros); fp32 atomics, so sums agree to rounding only. */
GLNN_API int glnn_spmm_csr_scatter_f32(const void* indptr, int indptr64, const int32_t* indices,
                              const float* dY, int64_t lddy, const float* scale, float* dX, int64_t lddx,
                              int64_t n_dst, int d, int self_add, glnn_stream_t stream);

/* Neighbour sampling for `train_sage` (dgl MultiLayerNeighborSampler + NodeDataLoader,
 * train_and_eval.py:179-190) on the device.  glnn_sample_count: counts[i] = min(in_degree(seeds[i]),
 * fanout) (fanout < 0: the whole neighbourhood).  glnn_sample_neighbors: with out_ptr = exclusive
 * prefix sum of the counts, writes for every seed its sampled in-edge SOURCES (global ids) to
 * out_src[out_ptr[i] ...]: all of them when in_degree <= fanout, else `fanout` distinct edges drawn
 * uniformly without replacement (Floyd's algorithm on a counter-based hash of (rng_seed, seed node)),
 * in CSR order.  fanout <= 64.  glnn_block_mark sets flag[src] = 1 for every sampled source and
 * glnn_block_relabel rewrites global ids to block-local ids through map[node] (dst nodes first,
 * models.py:105-109). */
GLNN_API int glnn_sample_count(const void* indptr, int indptr64, const int64_t* seeds, int64_t m, int fanout,
                      int64_t* counts, glnn_stream_t stream);
GLNN_API int glnn_sample_neighbors(const void* indptr, int indptr64, const int32_t* indices,
                          const int64_t* seeds, int64_t m, int fanout, uint64_t rng_seed,
                          const int64_t* out_ptr, int32_t* out_src, glnn_stream_t stream);
GLNN_API int glnn_block_mark(const int32_t* src, int64_t total, uint8_t* flag, glnn_stream_t stream);
GLNN_API int glnn_block_relabel(int32_t* src, int64_t total, const int32_t* map, glnn_stream_t stream);

/* Graph construction on the device (what DGL does on the host for dgl.graph((src, dst)) /
 * g.create_formats_(), dataloader.py:78,105 and train_and_eval.py:178, and for g.subgraph(idx_obs),
 * train_and_eval.py:324).
 * glnn_csr_from_coo: edge list -> CSR over destinations, int32, STABLE (the edges of a row keep their
 * input order; multi-edges kept): degree counts + exclusive scan + a least-significant-digit radix
 * sort of (dst, src) pairs, 8 bits per pass.  src / dst are int64 (idx64 = 1) or int32 device arrays;
 * indptr [n_nodes + 1]; indices [n_edges]; out_deg [n_nodes] int64 or NULL.  *status (device int32)
 * receives the number of edges with an id outside [0, n_nodes) -- the caller must treat non-zero as an
 * error (such edges are stored as 0 -> 0 to keep the kernels in bounds).  Fewer than 2^31 edges.
 * glnn_csr_subgraph: node-induced subgraph; relabel[v] = new id of old node v or -1.  Row
 * relabel[v] of the result holds the relabelled kept sources of old row v in their old order.
 * new_indptr [n_new + 1]; with new_indices == NULL and new_out_deg == NULL only new_indptr is computed
 * (read new_indptr[n_new], allocate, call again); new_out_deg [n_new] int64 or NULL.
 * Workspaces: 256-byte aligned device memory; glnn_csr_build_workspace_bytes(n_nodes, n_edges) for
 * from_coo, 12 * (n_new + 1) + 1024 bytes suffice for subgraph. */
GLNN_API size_t glnn_csr_build_workspace_bytes(int64_t n_nodes, int64_t n_edges);
GLNN_API int glnn_csr_from_coo(const void* src, const void* dst, int idx64, int64_t n_edges,
                               int64_t n_nodes, int32_t* indptr, int32_t* indices, int64_t* out_deg,
                               int32_t* status, void* workspace, size_t ws_bytes, glnn_stream_t stream);
GLNN_API int glnn_csr_subgraph(const void* indptr, int indptr64, const int32_t* indices, int64_t n_nodes,
                               const int32_t* relabel, int64_t n_new, int32_t* new_indptr,
                               int32_t* new_indices, int64_t* new_out_deg, void* workspace,
                               size_t ws_bytes, glnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Host-buffer entry point (what an out-of-process / non-torch caller binds; used for the e2e
 * benchmark leg).  SAGE("gcn") eval forward = SAGE.inference + log_softmax: copies the CSR graph,
 * features and weights from HOST memory, runs glnn_sage_forward on the device and copies the
 * [n, label_dim] log-probabilities back.  Synchronous; allocates and frees its own device memory.
 * Weights follow the state_dict layout: per layer W [out,in], b [out]; per hidden layer optional
 * BN (gamma, beta, running_mean, running_var), NULL when norm_type="none".
 */
typedef struct glnn_sage_layer_host {
  const float* weight; /* [d_out, d_in] */
  const float* bias;   /* [d_out] */
  const float* bn_gamma;
  const float* bn_beta;
  const float* bn_mean;
  const float* bn_var;
  int32_t d_in, d_out;
} glnn_sage_layer_host;

GLNN_API int glnn_sage_inference_host(const int64_t* indptr_host, const int32_t* indices_host, int64_t n,
                             const float* feats_host, const glnn_sage_layer_host* layers,
                             int num_layers, float bn_eps, float* out_logprob_host);

#ifdef __cplusplus
}
#endif
#endif /* GLNN_B200_H_ */
