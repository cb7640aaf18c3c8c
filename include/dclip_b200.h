/* dclip_b200 — C ABI of the B200-native (sm_100a) DCLIP hot path.
 *
 * The reference (ChuckDanz/DCLIP) is pure Python and has no FFI layer: its "boundary" is a
 * set of Python callables that the training / eval scripts import by name.
 * This header is the C-ABI those callables bind to in the B200 build (via ctypes, see
 * dclip_b200/_lib.py and INTEGRATION.md).  Each entry point cites the reference code it
 * replaces (paths relative to the reference tree).
 *
 * Conventions
 *  - plain pointers and sizes only; every data pointer is a DEVICE pointer unless noted
 *  - the library never allocates or frees tensor memory: inputs, outputs and workspace are
 *    caller-owned; `*_workspace_bytes()` tells how much scratch an op needs
 *  - every op only enqueues work on `stream` (a cudaStream_t passed as void*); no device sync
 *  - returns DCLIP_OK (0) or a negative dclip_status; message via dclip_last_error()
 *  - no CPU fallback: on a machine without an sm_100 device every compute call fails
 *  - row-major, contiguous tensors; `d` (embedding dim) must be a multiple of 8
 */
#ifndef DCLIP_B200_H
#define DCLIP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCLIP_B200_VERSION 100 /* major*10000 + minor*100 + patch */

typedef enum {
    DCLIP_OK = 0,
    DCLIP_ERR_INVALID = -1,     /* bad shape / null pointer / misalignment */
    DCLIP_ERR_UNSUPPORTED = -2, /* combination not implemented (e.g. d too large for a path) */
    DCLIP_ERR_WORKSPACE = -3,   /* workspace too small */
    DCLIP_ERR_CUDA = -4,        /* CUDA runtime error (message has cudaGetErrorString) */
    DCLIP_ERR_ARCH = -5         /* tensor path requested on a non-sm_100 device */
} dclip_status;

typedef enum { DCLIP_F32 = 0, DCLIP_BF16 = 1, DCLIP_F16 = 2 } dclip_dtype;

/* Which arithmetic a GEMM-shaped op runs in.
 *  FP32   : CUDA-core FFMA, fp32 end to end (1e-5 parity tier; bit-exact retrieval contract)
 *  TENSOR : tcgen05 / TMEM tiles fed by TMA, fp16 operands (normalised rows, |x|<=1) or the
 *           input 16-bit dtype, fp32 accumulation (1e-3 parity tier); retrieval re-scores
 *           candidates with the FP32 contract so indices stay bit-exact
 *  AUTO   : TENSOR for 16-bit inputs on large problems, FP32 otherwise                      */
typedef enum { DCLIP_PATH_AUTO = 0, DCLIP_PATH_FP32 = 1, DCLIP_PATH_TENSOR = 2 } dclip_path;

int dclip_version(void);
/* Copies the calling thread's last error message (NUL-terminated) into buf. */
int dclip_last_error(char* buf, size_t buf_len);
/* 1 if device `device` is compute capability 10.x (tcgen05 path usable). */
int dclip_device_supports_tensor_path(int device);

/* ------------------------------------------------------------------------------------------
 * Distillation objective   (replaces compute_contrastive_loss
 *   training/CLIP_image_distillation.py:532-562 == training/train_contrastive_teacher.py:251-261,
 *   cosine_distillation_loss training/CLIP_image_distillation.py:564-576 and the combine in
 *   training_step :594-634)
 *
 *   u_hat = s_img / max(|s_img|, 1e-12), v_hat likewise;  S = u_hat v_hat^T / tau
 *   L_con = 1/(2B) sum_i (lse_row_i + lse_col_i - 2 S_ii)
 *   L_img = mean_i (1 - u_hat_i . t_hat_img_i),  L_txt likewise
 *   L = w_img L_img + w_txt L_txt + w_con L_con ; gradients w.r.t. s_img, s_txt only.
 *
 * The B x B logits are never written to memory: forward keeps per-row / per-column
 * log-sum-exp vectors, backward recomputes score tiles.
 * ------------------------------------------------------------------------------------------ */

/* Single-GPU, whole objective, forward + backward in one call.
 *  s_img, s_txt : [B, d] student embeddings, dtype `dtype`
 *  t_img, t_txt : [B, d] teacher embeddings (same dtype) or NULL (pure InfoNCE: the teacher
 *                 step's loss, train_contrastive_teacher.py:357)
 *  losses       : float[4] = {total, L_img, L_txt, L_con}          (device)
 *  d_s_img/d_s_txt : [B, d] gradients in `dtype`, already multiplied by `grad_scale`
 *                 (the upstream d(total)); either both NULL (forward only) or both set
 *  stats        : optional float[2*B] = {lse_row[B], lse_col[B]} for callers that save them */
size_t dclip_distill_loss_workspace_bytes(int64_t B, int64_t d, int path);
int dclip_distill_loss_fwd_bwd(const void* s_img, const void* s_txt, const void* t_img, const void* t_txt,
                               int64_t B, int64_t d, int dtype, float temperature, float w_img, float w_txt,
                               float w_con, float grad_scale, float* losses, void* d_s_img, void* d_s_txt,
                               float* stats, void* workspace, size_t workspace_bytes, int path, void* stream);

/* Number of kernels of this library one step launches (single-GPU composite when sharded == 0, the
 * per-rank building blocks (1)-(5) when sharded != 0; sharded == 2: the fused phases with the peer exchange inside the
 * kernels, no push launches); used by bench.py for "gpu_launches". */
/* The same objective split where the TEACHER embeddings are first needed, for callers whose inputs arrive over
 * PCIe: _students (normalise, forward tiles, log-sum-exp, backward tiles) needs only the student rows and can run
 * while the teachers are still being copied; _teachers (cosine terms, the four losses, gradient tail) finishes
 * the step.  Both calls must get the SAME workspace (dclip_distill_loss_workspace_bytes), B, d, dtype, temperature,
 * w_con and path; results equal dclip_distill_loss_fwd_bwd up to the rounding of one reciprocal. */
int dclip_distill_loss_students(const void* s_img, const void* s_txt, int64_t B, int64_t d, int dtype, float temperature, float w_con,
                                void* workspace, size_t workspace_bytes, int path, void* stream);
int dclip_distill_loss_teachers(const void* s_img, const void* s_txt, const void* t_img, const void* t_txt, int64_t B, int64_t d,
                                int dtype, float temperature, float w_img, float w_txt, float w_con, float grad_scale, float* losses,
                                void* d_s_img, void* d_s_txt, void* workspace, size_t workspace_bytes, int path, void* stream);
int dclip_distill_loss_launch_count(int64_t m, int64_t n, int64_t d, int path, int sharded);
/* 1 when the tcgen05 tensor path of the loss is compiled in and validated for use. */
int dclip_tensor_path_ready(void);
/* Resolves DCLIP_PATH_AUTO for the objective (the single place it is decided): DCLIP_PATH_TENSOR for 16-bit
 * inputs (`dtype` != DCLIP_F32) when B >= 128 (global batch), d % 128 == 0, 128 <= d <= 768, temperature >= 0.025
 * and the current device is sm_100; DCLIP_PATH_FP32 otherwise.  An explicit path is returned unchanged.  The host layer resolves once
 * and hands the result to the workspace query and to the single-GPU / sharded entry points alike. */
int dclip_resolve_loss_path(int path, int dtype, int64_t B, int64_t d, float temperature);

/* Test / debug hook: 1 if a tensor-path kernel hit a bounded-wait timeout since the last call
 * (clears the flag).  Synchronises `stream`; never called on the product path. */
int dclip_debug_pipeline_error(void* stream);

/* Sharded (data-parallel) building blocks — SURVEY.md §8(e).  Rank r owns `m` rows starting at
 * global row `row_offset` of a global batch of `n` rows.  The host all-gathers the normalised
 * operands between (1) and (2) and the two lse vectors between (2) and (3); no [B, d]
 * reduce-scatter is needed because each rank recomputes the column block it owns.
 *
 * (1) normalise the local student rows into GEMM operands and compute the row-local cosine
 *     terms.  x_hat: [m, d] in `op_dtype` (DCLIP_F32 for the FP32 path, DCLIP_F16 for TENSOR);
 *     inv_norm: float[m] = 1 / max(|x|, 1e-12);  cos: float[m] = x_hat . t_hat (0 if teacher NULL). */
int dclip_loss_prep(const void* x, const void* teacher, int64_t m, int64_t d, int dtype, void* x_hat,
                    int op_dtype, float* inv_norm, float* cos, void* stream);
/* (2) lse[i] = log sum_j exp(x_hat_i . y_hat_j / tau)  over all n rows of y_hat;
 *     diag[i] = x_hat_i . y_hat_{row_offset+i} / tau  (NULL to skip). */
size_t dclip_infonce_rowlse_workspace_bytes(int64_t m, int64_t n, int64_t d, int path);
int dclip_infonce_rowlse(const void* x_hat, int64_t m, const void* y_hat, int64_t n, int64_t d, int op_dtype,
                         float temperature, int64_t row_offset, float* lse, float* diag, void* workspace,
                         size_t workspace_bytes, int path, void* stream);
/* (3) dx_hat[i,:] = coef/tau * sum_j (exp(S_ij - lse_x[i]) + exp(S_ij - lse_y[j]) - 2 [j == row_offset+i]) y_hat[j,:]
 *     (fp32 [m, d]); lse_x: float[m] (local rows), lse_y: float[n] (all rows of the other side). */
size_t dclip_infonce_grad_workspace_bytes(int64_t m, int64_t n, int64_t d, int path);
int dclip_infonce_grad(const void* x_hat, int64_t m, const void* y_hat, int64_t n, int64_t d, int op_dtype,
                       float temperature, int64_t row_offset, const float* lse_x, const float* lse_y, float coef,
                       float* dx_hat, void* workspace, size_t workspace_bytes, int path, void* stream);
/* (3b) both directions of (3) in ONE launch for the single-GPU case (m == n == B, row_offset 0):
 *      du_hat from (u_hat, v_hat, lse_u, lse_v) and dv_hat from (v_hat, u_hat, lse_v, lse_u). */
size_t dclip_infonce_grad_pair_workspace_bytes(int64_t B, int64_t d, int path);
int dclip_infonce_grad_pair(const void* u_hat, const void* v_hat, int64_t B, int64_t d, int op_dtype, float temperature,
                            const float* lse_u, const float* lse_v, float coef, float* du_hat, float* dv_hat,
                            void* workspace, size_t workspace_bytes, int path, void* stream);
/* (3c) (3b) in two stages through a MATERIALISED coefficient matrix (tensor path; what the single-GPU composite runs
 *      for B > 128): the score tiles are recomputed once instead of once per side.
 *      stage 1: coeffs[i][j] = 2^13 * (exp(S_ij - lse_u[i]) + exp(S_ij - lse_v[j]) - 2 [i == j]) in float16, row stride
 *               dclip_infonce_coeffs_ld(B), dclip_infonce_coeffs_bytes(B) bytes; dots (dclip_infonce_coeffs_dot_bytes)
 *               receives partial row / column sums of coeffs * S = the radial components u_hat_i . du_hat_i, v_hat_j . dv_hat_j.
 *      stage 2: du_hat = coef/tau * 2^-13 * coeffs . v_hat,  dv_hat = coef/tau * 2^-13 * coeffs^T . u_hat  (fp32 [B, d]) as one
 *               launch of a CTA-pair GEMM; with `tail` its epilogue applies step (4) to both sides instead and writes
 *               d_s_img / d_s_txt in `dtype` (du_hat / dv_hat may then be NULL).  tinv_* = 1 / max(|t_i|, 1e-12). */
typedef struct dclip_grad_tail {
    const void* s_img; const void* s_txt; const void* t_img; const void* t_txt;       /* raw rows [B, d] in `dtype`; teachers may be NULL */
    const float* inv_u; const float* inv_v; const float* cos_img; const float* cos_txt; const float* tinv_img; const float* tinv_txt;
    float w_img_over_B, w_txt_over_B, grad_scale;
    int dtype;
    void* d_s_img; void* d_s_txt;
} dclip_grad_tail;
size_t dclip_infonce_coeffs_bytes(int64_t B);
int64_t dclip_infonce_coeffs_ld(int64_t B);
size_t dclip_infonce_coeffs_dot_bytes(int64_t B);
int dclip_infonce_coeffs(const void* u_hat, const void* v_hat, int64_t B, int64_t d, float temperature, const float* lse_u,
                         const float* lse_v, void* coeffs, float* dots, void* stream);
int dclip_infonce_grad_from_coeffs(const void* coeffs, const float* dots, const void* u_hat, const void* v_hat, int64_t B, int64_t d,
                                   float temperature, float coef, float* du_hat, float* dv_hat, const dclip_grad_tail* tail, void* stream);
/* (4) row-local tail: adds the cosine-distillation gradient (-w_cos/B_global * t_hat), applies the
 *     normalise backward  dx = (g - x_hat (x_hat . g)) * inv_norm  and writes d_x in `dtype`.
 *     dx_hat may be NULL (cosine term only).  teacher may be NULL. */
int dclip_loss_grad_tail(const void* x, const void* teacher, const float* inv_norm, const float* dx_hat, int64_t m,
                         int64_t d, int dtype, float w_cos_over_B, float grad_scale, void* d_x, void* stream);
/* (5) local partial sums of the loss terms (host all-reduces them):
 *     out[0] = sum_i (lse_u[i] + lse_v[i] - 2 diag[i]),  out[1] = sum_i (1 - cos_img[i]),
 *     out[2] = sum_i (1 - cos_txt[i]).  cos_* may be NULL.  Deterministic (fixed-order) sums. */
int dclip_loss_partial_sums(const float* lse_u, const float* lse_v, const float* diag, const float* cos_img,
                            const float* cos_txt, int64_t m, float* out3, void* stream);

/* Fused form of (1)-(5) for the tensor path: three calls per rank and step, 7 kernel launches, with the
 * host's all-gathers in between.  u_all / v_all: float16 [n, d] gather buffers (this rank writes rows
 * [row_offset, row_offset + m) in _prep, the host all-gathers them in place).  stat: float [world][2m + 4]
 * gather buffer; _fwd fills this rank's row {lse_u[m], lse_v[m], 3 raw loss sums}, the host all-gathers it,
 * _bwd reads every rank's row (blocked lse lookup inside the kernel, global losses from the raw sums).
 *
 * `peers` (NULL or world <= 1: none): the exchanges happen INSIDE these kernels over NVLink peer memory and the host does
 * nothing between the calls.  u_all / v_all / stat then live in the rank's symmetric buffer (dclip_p2p_alloc; same layout on
 * every rank; bases[r] = rank r's buffer as mapped here, bases[rank] = the local one; flag block at flags_offset):
 *   _prep  waits for every peer's "step done" credit, writes its normalised rows into EVERY rank's u_all / v_all and signals;
 *   _fwd   waits (TMA producers) for every peer's rows, writes its statistics row into every rank's stat and signals;
 *   _bwd   waits for every peer's statistics, and its gradient tail signals "step done" (gradients are mandatory here).
 * 16-bit inputs with d <= 1024 only.  A peer that never arrives is fatal (see the push / wait kernels below). */
typedef struct dclip_peers { void* const* bases; int world; int rank; size_t flags_offset; } dclip_peers;
size_t dclip_sharded_loss_workspace_bytes(int64_t m, int64_t n, int64_t d);
int dclip_sharded_loss_prep(const void* s_img, const void* s_txt, const void* t_img, const void* t_txt, int64_t m, int64_t d,
                            int dtype, float temperature, void* u_all, void* v_all, int64_t row_offset, float* inv_u,
                            float* inv_v, float* cos_img, float* cos_txt, float* diag, const dclip_peers* peers, void* stream);
int dclip_sharded_loss_fwd(const void* u_all, const void* v_all, int64_t m, int64_t n, int64_t d, float temperature,
                           int64_t row_offset, const float* cos_img, const float* cos_txt, const float* diag, float* stat_row,
                           void* workspace, size_t workspace_bytes, const dclip_peers* peers, void* stream);
int dclip_sharded_loss_bwd(const void* s_img, const void* s_txt, const void* t_img, const void* t_txt, const void* u_all,
                           const void* v_all, int64_t m, int64_t n, int64_t d, int dtype, float temperature, int64_t row_offset,
                           int rank, int world, const float* stat_all, const float* inv_u, const float* inv_v, float w_img,
                           float w_txt, float w_con, float grad_scale, float* losses, void* d_s_img, void* d_s_txt,
                           void* workspace, size_t workspace_bytes, const dclip_peers* peers, void* stream);

/* ------------------------------------------------------------------------------------------
 * Push-based exchange over NVLink peer memory (replaces NCCL's all-gathers in the data-parallel objective;
 * "all-gather of the normalised operands", "all-gather of the LSE vectors").
 * Every rank allocates ONE symmetric buffer (same layout everywhere) that also holds a flag block of
 * dclip_p2p_flags_bytes() at `flags_offset`; the 64-byte handles are exchanged by the host (torch.distributed)
 * and opened in every peer process.  peer_bases[r] = address of rank r's buffer in THIS process
 * (peer_bases[rank] = the local allocation).
 *   push:  optional wait on `wait_kind` (all peers), copy `nranges` byte ranges of the local buffer to the same
 *          offsets of every peer, then bump arrive[signal_kind][rank] on every peer (nranges = 0: signal only);
 *          wait_after = 1 also performs the wait below for `signal_kind` inside the same launch (all-gather)
 *   wait:  block the stream until every peer has pushed `kind` once more than this rank has consumed so far
 * Kinds are small integers (0..3) chosen by the caller.  Everything is asynchronous on `stream` and can be
 * captured in a CUDA graph.  A peer that never arrives is FATAL, as with NCCL's watchdog: after 120 s the waiting kernel sets the
 * error flag (dclip_debug_pipeline_error) and traps, so the next CUDA call of the process fails.
 */
size_t dclip_p2p_flags_bytes(void);
int dclip_p2p_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64);
int dclip_p2p_init_flags(void* base, size_t flags_offset, int credit_kind);
int dclip_p2p_open(const unsigned char* handle64, void** dev_ptr);
int dclip_p2p_close(void* dev_ptr);
int dclip_p2p_free(void* dev_ptr);
int dclip_p2p_push(void* const* peer_bases, int world, int rank, const size_t* offsets, const size_t* bytes, int nranges,
                   size_t flags_offset, int wait_kind, int signal_kind, int wait_after, void* stream);
int dclip_p2p_wait(void* local_base, size_t flags_offset, int world, int rank, int kind, void* stream);

/* ------------------------------------------------------------------------------------------
 * Retrieval   (replaces the similarity build + calculate_retrieval_metrics,
 *   eval_scripts/flickr30k_eval.py:243-266 and :16-88, and the faiss IndexFlatIP add/search
 *   used by training/compute_faiss.py:27,79 and training/image_tokenizer.py:262)
 *
 * Score contract (bit-exact with oracle/retrieval_oracle.c):
 *     score(q, g) = fp32 fmaf chain over k = 0..d-1;   order = (-score, +index).
 * The score matrix is never materialised.
 * ------------------------------------------------------------------------------------------ */

/* x_hat = x / sqrt(fmaf-chain sum of squares)   (numpy semantics: no eps).  out: float[n, d]. */
int dclip_normalize_rows(const void* x, int64_t n, int64_t d, int dtype, float* out, void* stream);

/* out[p] = score(q[qidx[p]], g[gidx[p]])  (qidx NULL -> p).  q, g: float32. */
int dclip_pair_scores(const float* q, const float* g, int64_t d, const int64_t* qidx, const int64_t* gidx,
                      int64_t npairs, float* out, void* stream);

/* Partial ranks over one gallery shard g[0..ng) whose rows have global indices
 * g_index_offset + j:   rank[p] = #{ j : s_pj > ts[p]  or (s_pj == ts[p] and g_index_offset + j < tgt[p]) }
 * with s_pj = score(q[qidx[p]], g[j]), ts = target scores from dclip_pair_scores, tgt = global
 * target index.  Ranks of shards add up (all-reduce sum) to the stable-argsort rank.
 * path = DCLIP_PATH_TENSOR (sm_100, d % 8 == 0, 64 <= d <= 1024): the score tiles run on tcgen05 with fp16 copies of the rows;
 * they differ from the fp32 chain by at most 1.1e-3 |q|max |g|max, so every entry outside that band around ts[p] is decided by the
 * tile and the few inside are re-scored with the exact chain - same ranks, bit for bit; data that defeats the band (its list
 * overflows) is recomputed by the exact kernel inside the same call. */
size_t dclip_rank_counts_workspace_bytes(int64_t npairs, int64_t ng, int64_t d, int path);
int dclip_rank_counts(const float* q, const float* g, int64_t ng, int64_t d, const int64_t* qidx,
                      const float* target_score, const int64_t* target_index, int64_t npairs,
                      int64_t g_index_offset, int32_t* rank, void* workspace, size_t workspace_bytes, int path,
                      void* stream);

/* rank_counts for query rows with SEVERAL targets each (the image -> text direction of calculate_retrieval_metrics,
 * eval_scripts/flickr30k_eval.py:48-64): the pairs of query row r are target_ptr[r] .. target_ptr[r + 1] (CSR over
 * target_score / target_index / rank).  One score tile per unique query row, compared against all of its targets:
 * same contract, same result as dclip_rank_counts with a gather index, a fraction of the FLOPs.
 * max_targets (host-known maximum of targets per row) must be <= 8. */
int dclip_rank_counts_multi(const float* q, const float* g, int64_t nq, int64_t ng, int64_t d, const int64_t* target_ptr,
                            const float* target_score, const int64_t* target_index, int64_t npairs, int max_targets,
                            int64_t g_index_offset, int32_t* rank, void* stream);
/* Compatibility form of calculate_retrieval_metrics (flickr30k_eval.py:16), which receives a
 * MATERIALISED similarity matrix sim[n_rows, n_cols] (float32 if is_f64 == 0 else float64, the
 * reference stores fp32 products in a float64 buffer, :255).  Pair p ranks element
 * target_index[p] within row (along_rows = 1) or column (along_rows = 0) line_index[p]:
 *   rank[p] = #{ j : s_j > s_t  or (s_j == s_t and j < t) }. */
int dclip_matrix_ranks(const void* sim, int is_f64, int64_t n_rows, int64_t n_cols, const int64_t* line_index,
                       const int64_t* target_index, int64_t npairs, int along_rows, int32_t* rank, void* stream);

/* R@1, R@5, R@10 and mean reciprocal rank ("MAP" in the reference, :70-71) of `n` ranks.
 * group (int64[n], values in [0, n_groups)) or NULL: when given, the best (minimum) rank per
 * group is scored and empty groups are skipped (i2t direction, :48-64).
 * out: double[4] on the device.  workspace: >= 4 * n_groups + 64 bytes when group != NULL. */
int dclip_recall_metrics(const int32_t* rank, const int64_t* group, int64_t n, int64_t n_groups, double* out4,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Exact inner-product top-k of every query against gallery shard g[0..ng):
 *   out_scores float[nq, k] descending, out_index int64[nq, k] = g_index_offset + row,
 *   ties by ascending index; unfilled slots (-inf, -1) as in faiss.  k <= 32.
 *   q: [nq, d] q_dtype, g: [ng, d] g_dtype (DCLIP_F32 or DCLIP_F16). */
size_t dclip_topk_ip_workspace_bytes(int64_t nq, int64_t ng, int64_t d, int64_t k, int q_dtype, int g_dtype,
                                     int path);
int dclip_topk_ip(const void* q, const void* g, int64_t nq, int64_t ng, int64_t d, int64_t k, int q_dtype,
                  int g_dtype, int64_t g_index_offset, float* out_scores, int64_t* out_index, void* workspace,
                  size_t workspace_bytes, int path, void* stream);
/* Tensor path of the exact top-k (tcgen05 candidate generation + exact re-scoring + certificate):
 *   q        [nq, d] queries in q_dtype
 *   g_exact  [ng, d] the stored gallery (g_dtype) that scores are DEFINED on (same contract as above)
 *   g_fp16   the same gallery in float16 for the tensor pass (== g_exact when g_dtype is DCLIP_F16)
 *   gmax     device float: max gallery row norm, from dclip_max_row_norm (cache it per index)
 *   flags    int32[nq]: 1 = the exactness certificate failed for that query (more than 32 near-ties):
 *            re-run it with dclip_topk_ip on the FP32 path; n_flagged: device int32 count.
 * Unflagged rows of out_scores / out_index are bit-identical to dclip_topk_ip(FP32).
 * Needs d % 8 == 0, 64 <= d <= 1024, k <= 16, ng < 2^26 per shard. */
int dclip_max_row_norm(const void* x, int64_t n, int64_t d, int dtype, float* out, void* stream);
size_t dclip_topk_ip_tc_workspace_bytes(int64_t nq, int64_t ng, int64_t d, int q_dtype);
int dclip_topk_ip_tc(const void* q, int q_dtype, const void* g_exact, int g_dtype, const void* g_fp16, int64_t nq,
                     int64_t ng, int64_t d, int64_t k, int64_t g_index_offset, const float* gmax, float* out_scores,
                     int64_t* out_index, int32_t* flags, int32_t* n_flagged, void* workspace, size_t workspace_bytes,
                     void* stream);

/* Merge `nlists` per-shard results ([nlists, nq, k] each) into one [nq, k] by (-score, +index). */
int dclip_topk_merge(const float* in_scores, const int64_t* in_index, int64_t nlists, int64_t nq, int64_t k,
                     float* out_scores, int64_t* out_index, void* stream);

/* ------------------------------------------------------------------------------------------
 * Meta-teacher cross-attention   (replaces CrossModalAttention.forward,
 *   training/patch_text_aggregation.py:13-46 = 2 x nn.MultiheadAttention + residual + LayerNorm,
 *   PatchTextAggregation.aggregation :243-265, the combine / NaN guard :640-656, and the autograd
 *   backward of the teacher step, training/train_contrastive_teacher.py:347-362)
 *
 * text [B, T, d], image [B, R, d] (batch-first, zero-padded rows are attended: there is no mask in
 * the reference).  params / grads: 12 pointers in the reference module's state_dict order
 *   0 text_to_image.in_proj_weight [3d,d]   1 .in_proj_bias [3d]   2 .out_proj.weight [d,d]   3 .out_proj.bias [d]
 *   4 image_to_text.in_proj_weight          5 .in_proj_bias        6 .out_proj.weight         7 .out_proj.bias
 *   8 norm_text.weight [d]   9 norm_text.bias   10 norm_image.weight   11 norm_image.bias
 * all in `dtype`.  `saved` keeps the fp32 intermediates the backward re-reads (Q, K|V, attention
 * output, pre-LayerNorm sums, LayerNorm statistics); attention probabilities are recomputed.
 * ------------------------------------------------------------------------------------------ */
size_t dclip_xattn_saved_bytes(int64_t B, int64_t T, int64_t R, int64_t d, int dtype);
size_t dclip_xattn_bwd_scratch_bytes(int64_t B, int64_t T, int64_t R, int64_t d, int dtype);
int dclip_xattn_fwd(const void* text, const void* image, int64_t B, int64_t T, int64_t R, int64_t d, int heads, int dtype,
                    const void* const* params, void* text_out, void* image_out, void* saved, size_t saved_bytes,
                    void* stream);
/* d_text / d_image may both be NULL (the teacher step needs parameter gradients only). */
int dclip_xattn_bwd(const void* text, const void* image, int64_t B, int64_t T, int64_t R, int64_t d, int heads, int dtype,
                    const void* const* params, const void* saved, const void* d_text_out, const void* d_image_out,
                    void* const* grads, void* d_text, void* d_image, void* scratch, size_t scratch_bytes, void* stream);

/* One whole step for inputs in PINNED HOST memory, issued by ONE call (the per-step Python between the pieces - stream contexts,
 * event objects, four tensor copies - was 56 us of a 0.56 ms step): waits for ev_done of the previous step on the copy stream, copies
 * the four [B, d] inputs to d_in[0..3] there (students first; ev_students, ev_teacher_img, ev_teachers recorded behind the student
 * pair, the image teacher and the text teacher), runs dclip_distill_loss_students on the compute stream behind ev_students (the
 * contrastive part overlaps the teachers' copies), the image side's gradient tail behind ev_teacher_img and the text side's behind
 * ev_teachers; the last block of that launch finishes the four losses and stores them into h_losses - PINNED host memory of 8
 * floats, 16-byte aligned, written by the device through its mapped address: [0..3] = {total, img, txt, con}, [4] = a sequence word
 * stored last, which is what the call waits for (the one host wait of the reference's `loss.item()`,
 * training/train_contrastive_teacher.py:364; no D2H copy, no stream synchronisation).  d_losses receives the same four values.
 * Without gradients (d_s_img == d_s_txt == NULL) the teachers phase is one dclip_distill_loss_teachers call followed by a 16-byte
 * D2H copy and a stream synchronisation.  The caller owns streams and events (created once, ev_done recorded once before the first
 * call); gradients stay on the device in d_s_img / d_s_txt. */
int dclip_host_distill_step(const void* h_s_img, const void* h_s_txt, const void* h_t_img, const void* h_t_txt, void* const* d_in,
                            int64_t B, int64_t d, int dtype, float temperature, float w_img, float w_txt, float w_con,
                            float grad_scale, float* d_losses, float* h_losses, void* d_s_img, void* d_s_txt, void* workspace,
                            size_t workspace_bytes, int path, void* compute_stream, void* copy_stream, void* ev_students,
                            void* ev_teacher_img, void* ev_teachers, void* ev_done);

/* out[i] = x[i] * (*scale), scale a float on the DEVICE: the upstream gradient of the scalar loss
 * (`loss.backward()`, training/train_contrastive_teacher.py:361; CLIP_image_distillation.py:634 through Lightning) applied to
 * the gradients that dclip_distill_loss_fwd_bwd already produced - no host read of the factor, graph-capturable.
 * n % 8 == 0; x / out of `dtype`, 16-byte aligned; out may alias x. */
int dclip_scale_grad(const void* x, int64_t n, int dtype, const float* scale, void* out, void* stream);

/* aggregation(attended [B, L, d], temperature = 2.0) -> [B, d]; weights: float[B, L] softmax weights
 * saved for the backward (may be NULL for inference). */
int dclip_aggregate_fwd(const void* x, int64_t B, int64_t L, int64_t d, int dtype, float temperature, void* out,
                        float* weights, void* stream);
int dclip_aggregate_bwd(const void* x, const float* weights, const void* d_out, int64_t B, int64_t L, int64_t d, int dtype,
                        float temperature, void* d_x, void* stream);

/* out = 0.5 a + 0.5 b; if any element is NaN/Inf the WHOLE tensor becomes zero (reference :649-651).
 * flag: int[1] device (1 = guard fired), kept by the caller for the backward.  No host sync. */
int dclip_combine_guard_fwd(const void* a, const void* b, int64_t n, int dtype, void* out, int* flag, void* stream);
int dclip_combine_guard_bwd(const void* d_out, int64_t n, int dtype, const int* flag, void* d_a, void* d_b, void* stream);

/* Fused tail of compute_global_embedding_batch (training/patch_text_aggregation.py:640-656): both aggregations
 * (:243-265), the 0.5 / 0.5 combine (:647) and the whole-tensor NaN / Inf guard (:649-651) in ONE launch, and the
 * backward of the same in one launch.  attended_text [B, T, d], attended_image [B, R, d], out [B, d]; d % 8 == 0,
 * d <= 1024.  w_text [B, T] / w_image [B, R]: the softmax weights (saved for the backward).  flags: int[B + 1]
 * (dclip_global_embedding_flags_bytes): per-element non-finite flags + a completion counter; device-side decision,
 * no host synchronisation.  zeroed_counter: optional device int that is ZERO on entry and is left zero by every launch
 * (a persistent word per stream owned by the caller); NULL: flags[B] is used and cleared by a memset on the stream. */
size_t dclip_global_embedding_flags_bytes(int64_t B);
int dclip_global_embedding_fwd(const void* attended_text, const void* attended_image, int64_t B, int64_t T, int64_t R,
                               int64_t d, int dtype, float temperature, void* out, float* w_text, float* w_image,
                               int* flags, int* zeroed_counter, void* stream);
int dclip_global_embedding_bwd(const void* attended_text, const void* attended_image, const float* w_text,
                               const float* w_image, const void* d_out, const int* flags, int64_t B, int64_t T,
                               int64_t R, int64_t d, int dtype, float temperature, void* d_attended_text,
                               void* d_attended_image, void* stream);

/* ------------------------------------------------------------------------------------------
 * Ragged batch assembly - the step immediately upstream of the cross-attention:
 *   the zero-pad / truncate / stack loops of training/patch_text_aggregation.py:555-620 and the token filter +
 *   per-token text_projection loop of training/text_tokenizer.py:196-213, as a varlen layout.
 * ------------------------------------------------------------------------------------------ */
/* Token filter of text_tokenizer.py:199-201: token i of sample b is kept iff input_ids[b,i] > 0, 0 < i < L - 1 and
 * attention_mask[b,i] == 1.  lengths[b] = kept tokens, positions[b, 0..lengths[b]) = their positions, in order. */
int dclip_token_select(const int64_t* input_ids, const int64_t* attention_mask, int64_t B, int64_t L, int* lengths,
                       int* positions, void* stream);
/* cu_seqlens[B + 1] = exclusive scan of max(lengths, min_len); *max_len (optional, device) = the batch maximum. */
int dclip_cu_seqlens(const int* lengths, int64_t B, int min_len, int* cu_seqlens, int* max_len, void* stream);
/* out[B, max_len, d]: row t of sample b = its t-th source row for t < len_b (truncated at max_len), EXACT zeros
 * beyond (the padding stays attended downstream: no mask exists in the reference).
 *   flat source  : src [total, d], cu_seqlens [B + 1]               (positions NULL; lengths optional)
 *   strided source: src [B, L_src, d], lengths [B], positions [B, L_src]   (the projected hidden states)
 *   fallback [B, d] (optional): used as row 0 of a sample with no rows (text_tokenizer.py:209-211)
 *   nan_guard: a sample holding a NaN / Inf is written as all zeros (patch_text_aggregation.py:542-544);
 *              sample_flags int[B] receives the flags. */
int dclip_pad_gather(const void* src, int dtype, const int* cu_seqlens, const int* lengths, const int* positions,
                     int64_t L_src, const void* fallback, int nan_guard, int* sample_flags, int64_t B, int64_t max_len,
                     int64_t d, void* out, void* stream);
/* out[n, m] = x[n, k] W[m, k]^T (+ bias[m], fp32): all token projections as one GEMM (text_tokenizer.py:202-207).
 * 16-bit dtype: tcgen05, fp32 accumulation; DCLIP_F32: exact fp32 GEMM. */
int dclip_linear_rows(const void* x, const void* W, const float* bias, int64_t n, int64_t k, int64_t m, int dtype,
                      void* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused multi-tensor Adam / AdamW + global-norm clip: torch.optim.Adam at
 *   training/train_contrastive_teacher.py:245-248, :360-362; AdamW + gradient_clip_val = 0.5 at
 *   training/CLIP_image_distillation.py:679-682, training/CLIP_image_distill_training.py:41-42.
 * `state` = dclip_optim_state_bytes() of device memory holding the step counter and the clip coefficient, so the
 * two launches are graph-capturable.  Moments are fp32; parameters / gradients share `dtype`.  Up to 64 tensors.
 * ------------------------------------------------------------------------------------------ */
size_t dclip_optim_state_bytes(void);
size_t dclip_optim_workspace_bytes(const int64_t* numel, int count);
int dclip_optim_init_state(void* state, void* stream);
/* clip_coef = min(1, max_norm / (||all grads||_2 + 1e-6)) into `state` (torch.nn.utils.clip_grad_norm_); the
 * gradients themselves are left untouched, dclip_adam_step applies the coefficient.  max_norm <= 0: only the norm. */
int dclip_grad_clip_coef(const void* const* grads, const int64_t* numel, int count, int dtype, float max_norm, void* state,
                         void* workspace, size_t workspace_bytes, void* stream);
/* One optimizer step over all tensors.  lr_dev (optional, device float) overrides lr (schedulers inside a graph).
 * decoupled_weight_decay: 1 = AdamW, 0 = Adam's L2 form.  Hyper-parameters are doubles: like torch's Python scalars
 * they are rounded to fp32 only where they meet the tensors (1 - beta in double first). */
int dclip_adam_step(void* const* params, const void* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                    const int64_t* numel, int count, int dtype, double lr, const float* lr_dev, double beta1, double beta2,
                    double eps, double weight_decay, int decoupled_weight_decay, void* state, void* stream);
/* Host read of {grad_norm, clip_coef, step} (synchronises `stream`; logging / tests). */
int dclip_optim_read_state(const void* state, float* grad_norm, float* clip_coef, int* step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DCLIP_B200_H */
