// Internal launch-argument structs shared by the loss kernels (FP32 and tensor paths).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "p2p.h"

namespace dclip {

// A "side" is one direction of the symmetric objective: side 0 = image rows against all text
// rows, side 1 = text rows against all image rows.  Kernels take both sides in one launch
// (blockIdx.z / .y selects) so the single-GPU step is 6 launches.
struct PrepSide { const void* x; const void* t; void* xhat; float* inv_norm; float* cos; float* t_inv; };   // t_inv: 1 / max(|t|, eps), optional
struct PrepArgs {
    PrepSide s[2]; int64_t m, d; float* diag; float inv_tau;   // diag: pair kernel only
    PeerCtx peers; int* error_flag;      // peers.world > 1 (register-resident pair kernel): the operand rows are also written into every peer's
                                         // buffer (after its "step done" credit) and the last block signals kP2PKindOperands
};

// cosine terms of the rows against the teacher, students' inverse norms given (phase 2 of the split objective)
struct CosSide { const void* x; const void* t; const float* inv_norm; float* cos; };
struct CosArgs { CosSide s[2]; int64_t m, d; };

struct LseSide {
    const float* x; const float* y;   // FP32 path operands (normalised rows)
    float* part_m; float* part_l;     // [nsplit, m] online-softmax partials
    float* lse; float* diag;          // outputs [m]
};
struct LseArgs {
    LseSide s[2];
    int64_t m, n, d;
    float temperature;
    int64_t row_offset;
    int nsplit, tiles_per_split;
};

struct GradSide { const float* x; const float* y; const float* lse_x; const float* lse_y; float* dx; };
struct GradArgs {
    GradSide s[2];
    int64_t m, n, d;
    float temperature;
    int64_t row_offset;
    float coef_over_tau;
};

struct TailSide { const void* x; const void* t; const float* inv_norm; const float* dx_hat; void* d_x;
                  float* cos_out; };      // optional: cos(x, t) of the row (the split host-buffer step: saves the separate cosine pass)
// Teachers phase of the split objective: the LAST block of the (last) tail launch finishes the four losses from the cosine terms the
// tail launches left (TailSide::cos_out) and the contrastive sum of the students phase - no k_loss_sums launch behind the tail - and,
// for the host-buffer step, also stores them into mapped pinned host memory followed by a sequence word (no D2H copy, no event wait).
struct TailSums {
    unsigned int* counter;               // blocks done; zero on entry (cleared by the students phase), left zero
    const float* con_sum;                // raw sum of lse_u + lse_v - 2 diag (students phase)
    const float* cos_img; const float* cos_txt;      // [m]
    float* losses;                       // device {total, img, txt, con}
    float* losses_host;                  // optional: device-visible address of a pinned host float[8]; [4] receives seq after the losses
    unsigned int seq;
    float w_img, w_txt, w_con;
};
struct TailArgs {
    TailSide s[2]; int64_t m, d; float w_cos_over_B[2]; float grad_scale;
    PeerCtx peers;                       // peers.world > 1: the last block tells every peer that this rank is done with the step's buffers
    TailSums sums;                       // sums.counter != nullptr: see TailSums
};

struct SumsArgs {
    const float* lse_u; const float* lse_v; const float* diag; const float* cos_img; const float* cos_txt;
    int64_t m;
    float* out3;      // raw partial sums (sharded API) or nullptr
    float* losses;    // finished {total, img, txt, con} (single-GPU composite) or nullptr
    float w_img, w_txt, w_con;
    unsigned int* zero_counter;   // optional: cleared by the thread that writes the sums (TailSums::counter of the teachers phase)
};

// ---- tensor path (loss_tc.cu) ----
struct TcSide {
    const float* lse_x;   // [m]   (bwd)
    const float* lse_y;   // [n]   (bwd); with lse_y_block > 0 element j lives at (j / block) * stride + j % block
    float* out;           // fwd: partial row sums [nsplit][m];  bwd: partial dx [nsplit][m][d]
    int64_t lse_y_block, lse_y_stride;   // blocked layout of the all-gathered per-rank statistics (sharded path)
};
struct TcArgs {
    TcSide s[2];
    int64_t m, n, d;
    float inv_tau;        // 1 / tau
    int64_t row_offset;   // global index of local row 0 (diagonal)
    int tiles_per_split;
    float out_scale;      // bwd: coef / tau * 2^-13
    int* error_flag;      // set to 1 on a pipeline timeout
    float* cand_s;        // retrieval mode: [nsplit][m][32] approximate scores (descending)
    int32_t* cand_i;      //                 [nsplit][m][32] gallery rows (-1 = empty)
    float* col_out;       // stream mode 1: partial column sums [ceil(m / 128)][n]
    void* g_out;          // stream mode 3: fp16 gradient coefficients G * 2^13, [ceil(m / 128) * 128][ldg]
    int64_t ldg;          //                row stride (elements), a multiple of 256 >= n
    // stream kernel: soft lockstep of concurrently running CTAs (keeps the Y tiles they share inside L2); null = off
    unsigned int* sync_ctr;   // [groups][sync_epochs], zeroed before the launch
    int sync_every, sync_group, sync_epochs;
    // stream mode 4 (Recall@K rank counting): exact target score / global target index per row, rank accumulator, band list
    const float* rk_thr; const int64_t* rk_tgt; int32_t* rk_rank; int2* rk_band; unsigned int* rk_ctr; unsigned int rk_cap;
    float rk_eps; int64_t rk_goff;
    const float* rk_norms;   // {max |q|, max |g|} (device): the band half-width is rk_eps * |q|max * |g|max
    PeerCtx peers;        // stream mode 0, peers.world > 1: the producers wait for every peer's operand rows before the first load
    int debug_skip;       // timing experiments only (results are garbage): 1 = no phase-1 MMAs, 2 = no phase-2 MMAs, 4 = no G stores
};


bool tc_loss_shape_ok(int64_t d, float temperature);
int tc_nsplit(int64_t m, int64_t n, int64_t d, int nsides);
int launch_infonce_bwd_tc(int nsides, const void* const x[2], const void* const y[2], const TcArgs& args, int nsplit,
                          cudaStream_t st);
int launch_tc_lse_combine(const float* part, int nsplit, int64_t m, float inv_tau, float* lse, cudaStream_t st);
int launch_tc_grad_combine(const float* part, int nsplit, int64_t m, int64_t d, float* dx, cudaStream_t st);
// both sides in one launch: part = [2][nsplit][m][d], dx0 / dx1 the two outputs
int launch_tc_grad_combine2(const float* part, int nsplit, int64_t m, int64_t d, float* dx0, float* dx1, cudaStream_t st);
int launch_tc_diag(const void* x, const void* y, int64_t m, int64_t n, int64_t d, float inv_tau, int64_t row_offset, float* diag,
                   cudaStream_t st);
int tc_check_error_flag(cudaStream_t st);
// single-GPU backward through the materialised fp16 coefficient matrix G (stream mode 3 + two GEMMs); g_buf = tc_gmat_bytes(B)
bool tc_gmat_ok(int64_t B, int64_t d);
size_t tc_gmat_bytes(int64_t B);
// Gradient tail fused into the GEMM epilogue (normalise backward + cosine-term gradient, k_loss_grad_tail's math): per side
// the raw student / teacher rows, the prep kernel's per-row vectors and the output gradient; null `tail` = fp32 dX_hat outputs.
struct GmatTailSide { const void* x; const void* t; const float* inv_norm; const float* cos; const float* t_inv; void* d_x; float w_cos_over_B; };
struct GmatTail { GmatTailSide s[2]; float grad_scale; int dtype; };
// workspace for the row / column sums of g * s: floats
size_t tc_gmat_dot_floats(int64_t B);
int64_t tc_gmat_ld(int64_t B);
int launch_gmat_coeffs(const void* u_hat, const void* v_hat, int64_t B, int64_t d, const float* lse_u, const float* lse_v, float inv_tau,
                       void* g_buf, float* dot_ws, cudaStream_t st);
int launch_gmat_gemms(const void* g_buf, const float* dot_ws, const void* u_hat, const void* v_hat, int64_t B, int64_t d, float out_scale,
                      float* du_hat, float* dv_hat, const GmatTail* tail, cudaStream_t st);
int launch_infonce_bwd_gmat(const void* u_hat, const void* v_hat, int64_t B, int64_t d, const float* lse_u, const float* lse_v, float inv_tau,
                            float out_scale, void* g_buf, float* dot_ws, float* du_hat, float* dv_hat, const GmatTail* tail, cudaStream_t st);
// streaming 128 x 256 score-tile kernel (stream_tc.cu): mode 0 = row sums, mode 2 = top-32 candidates
int stream_nsplit(int64_t m, int64_t n, int nsides);
int launch_stream_tc(int mode, int nsides, const void* const x[2], const void* const y[2], const TcArgs& args, int nsplit,
                     cudaStream_t st);
// Recall@K rank counting on the tensor path (retrieval_tc.cu)
size_t rank_counts_tc_workspace_bytes(int64_t npairs, int64_t ng, int64_t d);
int launch_rank_counts_tc(const float* q, const float* g, int64_t ng, int64_t d, const int64_t* qidx, const float* ts, const int64_t* tgt,
                          int64_t npairs, int64_t goff, int32_t* rank, void* workspace, size_t workspace_bytes, int** gate_out, cudaStream_t st);
// lockstep counters the streaming kernel wants for this problem (0 = lockstep not used); see stream_tc.cu
size_t stream_sync_bytes(int64_t m, int64_t n, int64_t d, int nsplit);

int launch_loss_prep(const PrepArgs& a, int nsides, int dtype, int op_dtype, cudaStream_t st);
int launch_cos_rows(const CosArgs& a, int dtype, cudaStream_t st);
// both sides of row i in one warp + diag[i] = u_hat_i . v_hat_i / tau from the rounded operands
int launch_loss_prep_pair(const PrepArgs& a, int dtype, int op_dtype, cudaStream_t st);
// tensor path: lse = 1/tau + ln(sum of partial row sums) for both sides + the loss sums, one launch
struct FinalizeArgs {
    const float* part_u; const float* part_v; int nsplit; int64_t m; float inv_tau;
    float* lse_u; float* lse_v; SumsArgs sums;
    int nsplit_v;         // partial count of part_v when it differs from part_u's (0 = nsplit): column-sum partials
    PeerCtx peers;        // peers.world > 1: lse_u / lse_v / out3 live in the symmetric buffer and are also written to every peer, then
                          // kP2PKindStats is signalled; the kernel also retires this step's operand / done waits (expected[] += 1)
};
int launch_tc_finalize(const FinalizeArgs& a, cudaStream_t st);
// global losses {total, img, txt, con} from the all-gathered per-rank raw sums stat[r * stride + off + 0..2]
// peers (optional, world > 1): wait for every peer's statistics first (kP2PKindStats) and retire that wait
int launch_sharded_losses(const float* stat, int world, int64_t stride, int64_t off, int64_t B, int has_img, int has_txt,
                          float w_img, float w_txt, float w_con, float* losses, cudaStream_t st, const PeerCtx* peers = nullptr);
int simt_rowlse_nsplit(int64_t m, int64_t n, int64_t d);
int launch_rowlse_simt(LseArgs a, int nsides, cudaStream_t st);
int launch_grad_simt(const GradArgs& a, int nsides, cudaStream_t st);
int launch_grad_tail(const TailArgs& a, int nsides, int dtype, cudaStream_t st);
int launch_loss_sums(const SumsArgs& a, cudaStream_t st);
// whole objective for B <= 64 in one launch (loss_small.cu): one cluster of 8 CTAs, fp32 arithmetic
bool small_loss_shape_ok(int64_t B, int64_t d);
int launch_distill_small(const void* s_img, const void* s_txt, const void* t_img, const void* t_txt, int64_t B, int64_t d, int dtype,
                         float temperature, float w_img, float w_txt, float w_con, float grad_scale, float* losses, void* d_img, void* d_txt,
                         float* stats, cudaStream_t st);

}  // namespace dclip
