// Grouped tcgen05 GEMM (gemm_tc.cu): host-side problem description.
#pragma once
#include <cuda_runtime.h>

namespace dclip {

constexpr int kTcGemmMax = 8;

// C[M,N] (fp32, ldc) = A'(M,K) . B'(K,N) + bias[n] + residual[m,n];  A, B are 16-bit (bf16 or fp16).
//   a_mn = 0: A stored [M, K] (lda = row stride in elements);  a_mn = 1: A stored [K, M]
//   b_mn = 0: B stored [N, K];                                  b_mn = 1: B stored [K, N]
// bias / residual are fp32 (kind 0) or of the operands' 16-bit type (kind 1); outputs: fp32 C and / or 16-bit C16.
struct TcGemmProblem {
    const void* A; const void* B; float* C; const void* bias; const void* residual; void* C16;
    int M, N, K, lda, ldb, ldc, ldr;
    int a_mn, b_mn, accumulate;
    int ldc16;      // row stride of C16 (0 = ldc); C may be null when only the 16-bit output is wanted
    int bias_kind, res_kind;
    float alpha;    // C = alpha * A'B' + ... (0 = 1)
    int ksplit;     // > 1: the contraction is cut into `ksplit` slices of whole 64-wide k-blocks, slice s writes its partial
    long long split_stride;   // product to C + s * split_stride (fp32 output only, no bias / residual / accumulate); the caller sums
    // Segmented contraction (fp32 operands split into 16-bit terms, xattn.cu): the logical K axis is `K / seg_len` segments of
    // seg_len (a multiple of 64) and segment s reads PHYSICAL segment (seg_map >> 4 s) & 15 of the stored operand, whose K extent
    // is k_phys - one stored term serves several term products.  seg_len = 0: plain contraction.
    int seg_len, k_phys;
    unsigned seg_map_a, seg_map_b;
};

int launch_gemm_tc(const TcGemmProblem* probs, int count, int dtype16, cudaStream_t st);
// number of slices a problem with contraction length K and a requested `ksplit` is actually cut into (the partials to sum)
int gemm_tc_slices(int K, int ksplit);
// CTA-pair kernel with 256 x 384 tiles (gemm_pair.cu): plain fp32 outputs only, at most 2 problems
bool gemm_pair_ok(const TcGemmProblem& q);
// Optional fused epilogue of the distillation gradients (normalise backward + cosine-term gradient, per problem):
//   out[i, :] = ((alpha acc[i, :] + ts_i t[i, :]) - x[i, :] inv_i dot_i) inv_i grad_scale,   ts_i = -w t_inv_i,
//   dot_i = alpha sum_q dot_part[q][i] - w cos_i   (0 when the norm was clamped);  x / t / out are [M, N] in `io_dtype`.
struct GemmPairTail {
    const void* x; const void* t; const float* inv_norm; const float* cos; const float* t_inv;
    const float* dot_part; int dot_count; long long dot_stride;
    void* out; float w_cos_over_B;
};
int launch_gemm_pair(const TcGemmProblem* probs, int count, int dtype16, cudaStream_t st, const GemmPairTail* tails = nullptr,
                     float grad_scale = 1.0f, int io_dtype = 0);

}  // namespace dclip
